"""C4 at full size through the tensor-core path with the per-role wait profile (MRB_MMA_PROF=1)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import multirate_b200 as mr, multirate_oracle as mo
N = 32
hLen, beta = mo.kaiserlength(0.05, samplerate=N); hLen = -(-hLen // N) * N
h = (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(np.float32)
for nch in (1024, 8192):
    x = torch.rand((nch, 65536), device="cuda")
    f = mr.FIRFilter(h, 0.918734, N, None, nchannels=nch, sample_dtype=np.float32)
    for _ in range(3):
        f.filt(x)
    torch.cuda.synchronize()
    f.set_timing(True)
    for _ in range(10):
        f.filt(x)
    torch.cuda.synchronize()
    ms = f.kernel_ms()
    print("nch=%d kernel=%s %.3f ms -> %.1f Gout/s" % (nch, f.last_kernel, ms, f._exact_count(65536) * nch / ms / 1e6), flush=True)
