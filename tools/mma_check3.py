"""Integer-ratio float32 workloads: tensor-core kernel (policy 0) against the CUDA-core fast paths (policy 2)."""
import os, sys
from fractions import Fraction
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import multirate_b200 as mr, multirate_oracle as mo
cases = [("c3b standard-128", Fraction(1, 1), mo.firdes(128, 0.25, 7.8562), 4096), ("c3a interp 4//1", Fraction(4, 1), mo.firdes(128, 0.125, 7.8562) * 4, 4096),
         ("xr32 rational 147//160", Fraction(147, 160), mo.firdes(3528, 0.5 / 147, 7.8562), 8192), ("standard-32", Fraction(1, 1), mo.firdes(32, 0.25, 7.8562), 4096)]
for name, ratio, h, nch in cases:
    h = h.astype(np.float32)
    x = torch.rand((nch, 65536), device="cuda")
    res = {}
    for pol in (0, 2):
        f = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=np.float32)
        f.set_kernel_policy(pol)
        for _ in range(3):
            y = f.filt(x)
        torch.cuda.synchronize()
        f.set_timing(True)
        for _ in range(10):
            f.filt(x)
        torch.cuda.synchronize()
        ms = f.kernel_ms()
        res[pol] = (f.last_kernel, ms, y)
        print("%-26s policy %d kernel %-24s %.3f ms -> %.1f Gout/s" % (name, pol, f.last_kernel, ms, f._exact_count(65536) * nch / ms / 1e6), flush=True)
    d = (res[0][2] - res[2][2]).abs().max().item() / res[2][2].abs().max().item()
    print("    tensor-core vs CUDA-core normalised difference %.3g" % d, flush=True)
