mkdir -p gpurun_out
out=gpurun_out/s4_sweep26.txt; : > $out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 -k "long_stream" 2>&1 | tail -3 >> $out
python - >> $out 2>&1 <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
from fractions import Fraction
import multirate_b200 as mr, multirate_oracle as mo
h = mo.firdes(24 * 147, 0.5 / 147, 7.8562).astype(np.float32)
n = 1 << 31
x = torch.view_as_complex(torch.rand((n, 2), device='cuda'))
torch.cuda.synchronize()
for it in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y = mr.filt_long_stream(h, Fraction(147, 160), x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('2^31-sample complex64 stream, 147//160: %d outputs in %.2f ms = %.1f Gout/s, %.2f TB/s algorithmic' % (y.shape[0], ms, y.shape[0] / ms / 1e6, (n + y.shape[0]) * 8 / ms / 1e9))
    del y
# spot check against the oracle at the far end of the stream: a piece that starts on a multiple of M (phase 0)
y = mr.filt_long_stream(h, Fraction(147, 160), x)
n0 = 160 * 13_000_000
k0 = n0 * 147 // 160
w = mo.filt(h, x[n0:n0 + 4000].cpu().numpy(), Fraction(147, 160))     # zero history: the first ~24 outputs differ
yy = y[k0:k0 + w.shape[0]].cpu().numpy()
print('far-end spot check', float(np.abs(w[40:] - yy[40:]).max() / np.abs(yy).max()))
PY
cat $out
