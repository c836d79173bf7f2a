"""Quick diagnostic of the tensor-core path (not a test): arbitrary / farrow float32 against the generic kernel and the
C oracle, a few shapes, prints errors and kernel times.  Run on a B200."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import multirate_b200 as mr
import multirate_oracle as mo

N = 32
hLen, beta = mo.kaiserlength(0.05, samplerate=N)
hLen = -(-hLen // N) * N
h = (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(np.float32)
rng = np.random.default_rng(1)


def nerr(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))


for po in (None, 4):
    for rate, nch, n in ((0.918734, 128, 20000), (0.918734, 1024, 65536), (1.37, 200, 30000), (0.6, 64, 30000)):
        x = rng.random((nch, n), dtype=np.float32)
        xd = torch.from_numpy(x).cuda()
        f = mr.FIRFilter(h, rate, N, po, nchannels=nch, sample_dtype=np.float32)
        g = mr.FIRFilter(h, rate, N, po, nchannels=nch, sample_dtype=np.float32)
        g.set_kernel_policy(2)
        r = mr.FIRFilter(h, rate, N, po, nchannels=nch, sample_dtype=np.float32)
        r.set_kernel_policy(1)
        y = f.filt(xd)
        torch.cuda.synchronize()
        k1 = f.last_kernel
        yg = g.filt(xd)
        yr = r.filt(xd)
        torch.cuda.synchronize()
        o = mo.FIRFilter(h, rate, N, po)
        w = o.filt(x[0])
        print("po=%s rate=%.4f nch=%d n=%d kernel=%s/%s  mma-vs-generic %.3g  table-vs-generic %.3g  mma-vs-oracle %.3g  generic-vs-oracle %.3g"
              % (po, rate, nch, n, k1, g.last_kernel, nerr(y.cpu().numpy(), yr.cpu().numpy()), nerr(yg.cpu().numpy(), yr.cpu().numpy()),
                 nerr(y[0].cpu().numpy(), w), nerr(yr[0].cpu().numpy(), w)), flush=True)
        # timing: 10 calls each
        for name, ff in (("auto", f), ("cuda-core", g)):
            ff.set_timing(True)
            for _ in range(10):
                ff.filt(xd)
            torch.cuda.synchronize()
            ms = ff.kernel_ms()
            outs = f._exact_count(n) * nch
            print("    %-9s %.3f ms per call  -> %.1f Gout/s" % (name, ms, outs / ms / 1e6), flush=True)
