"""Soak test of the mbarrier-synchronised kernels added in round 2 (k_decim8, the FP64 tensor-core table kernel): the same
stream of chunks twice through fresh filters, every chunk's output compared BIT FOR BIT between the two passes (a data race on
the ring / staging / tap-row buffers shows up as a differing chunk), plus the first chunks against the generic kernel."""
import os, sys
from fractions import Fraction
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import multirate_b200 as mr, multirate_oracle as mo
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = 32
hl, beta = mo.kaiserlength(0.05, samplerate=N); hl = -(-hl // N) * N
cases = {
    "c4a64": (0.918734, (mo.firdes(hl, 0.45, beta, samplerate=32) * N).astype(np.float64), 1024, (N,), torch.float64, "table_f64_dmma"),
    "c4f64": (0.918734, (mo.firdes(hl, 0.45, beta, samplerate=32) * N).astype(np.float64), 1024, (N, 4), torch.float64, "table_f64_dmma"),
    "c2": (Fraction(1, 8), mo.firdes(256, 0.5 / 8, 7.8562).astype(np.float32), 1024, (), torch.complex64, "decim8_c64"),
    "c5": (Fraction(147, 160), mo.firdes(3528, 0.5 / 147, 7.8562).astype(np.float32), 2048, (), torch.complex64, "mma_c64_split"),
    "x4ac": (0.918734, (mo.firdes(hl, 0.45, beta, samplerate=32) * N).astype(np.float32), 1024, (N,), torch.complex64, "mma_c64_g32"),
}
torch.manual_seed(7)
for name, (ratio, h, nch, extra, dt, want) in cases.items():
    xs = [torch.randn((nch, 65536), device="cuda", dtype=dt) for _ in range(3)]
    sums = []
    for rep in range(2):
        f = mr.FIRFilter(h, ratio, *extra, nchannels=nch, sample_dtype={torch.float64: np.float64, torch.complex64: np.complex64}[dt])
        cs = []
        for i in range(steps):
            y = f.filt(xs[i % 3])
            cs.append(torch.view_as_real(y).view(torch.int32).to(torch.int64).sum() if dt == torch.complex64 else y.view(torch.int64).sum())
        torch.cuda.synchronize()
        assert f.last_kernel == want, f.last_kernel
        sums.append(torch.stack(cs).cpu())
    bad = int((sums[0] != sums[1]).sum())
    print("%s: %d chunks x 2 passes on %s, chunks that differ bit for bit: %d" % (name, steps, want, bad), flush=True)
    assert bad == 0
print("soak ok")
