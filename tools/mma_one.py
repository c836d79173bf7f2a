"""One workload through the automatic dispatch, a few calls (for ncu): python tools/mma_one.py c3b|c4a|xr32|c3a [nch]"""
import os, sys
from fractions import Fraction
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import multirate_b200 as mr, multirate_oracle as mo
w = sys.argv[1]
dt = np.float64 if w.endswith("64") else np.float32
w = w[:-2] if w.endswith("64") else w
N = 32
hl, beta = mo.kaiserlength(0.05, samplerate=N); hl = -(-hl // N) * N
cfg = {"c3b": (Fraction(1, 1), mo.firdes(128, 0.25, 7.8562), 4096, ()), "c3a": (Fraction(4, 1), mo.firdes(128, 0.125, 7.8562) * 4, 4096, ()),
       "c2": (Fraction(1, 8), mo.firdes(256, 0.5 / 8, 7.8562), 1024, ()),
       "c5": (Fraction(147, 160), mo.firdes(3528, 0.5 / 147, 7.8562), 8192, ()),
       "xr32": (Fraction(147, 160), mo.firdes(3528, 0.5 / 147, 7.8562), 8192, ()),
       "c4a": (0.918734, mo.firdes(hl, 0.45, beta, samplerate=32) * N, 1024, (N,)), "c4f": (0.918734, mo.firdes(hl, 0.45, beta, samplerate=32) * N, 1024, (N, 4))}[w]
ratio, h, nch, extra = cfg
if len(sys.argv) > 2:
    nch = int(sys.argv[2])
x = torch.rand((nch, 65536), device="cuda", dtype=torch.float64 if dt is np.float64 else torch.float32)
if w in ("c2", "c5"):
    x = torch.complex(x, torch.rand((nch, 65536), device="cuda"))
    dt = np.complex64
f = mr.FIRFilter(h.astype(np.float32 if dt is np.complex64 else dt), ratio, *extra, nchannels=nch, sample_dtype=dt)
for _ in range(4):
    f.filt(x)
torch.cuda.synchronize()
f.set_timing(True)
for _ in range(6):
    f.filt(x)
torch.cuda.synchronize()
ms = f.kernel_ms()
print("%s nch=%d kernel=%s %.3f ms -> %.1f Gout/s" % (w, nch, f.last_kernel, ms, f._exact_count(65536) * nch / ms / 1e6), flush=True)
