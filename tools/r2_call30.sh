o=gpurun_out; mkdir -p $o
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "float64_integer_kinds or stream_kernel or table_kernel" 2>&1 | tail -8
timeout 100 python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
from fractions import Fraction
import multirate_b200 as mr
rng=np.random.default_rng(0)
for ratio,nt,nch in ((Fraction(147,160),3528,4096),(Fraction(4,1),128,2048),(Fraction(1,1),63,4096)):
    h=rng.standard_normal(nt)
    x=torch.randn((nch,65536),device='cuda',dtype=torch.float64)
    for pol in (0,2):
        f=mr.FIRFilter(h,ratio,nchannels=nch,sample_dtype=np.float64)
        if pol==2:
            import os
        for _ in range(3): f.filt(x)
        torch.cuda.synchronize(); f.set_timing(True)
        for _ in range(5): y=f.filt(x)
        torch.cuda.synchronize()
        ms=f.kernel_ms(); print(ratio,nt,nch,f.last_kernel,round(ms,3),'ms',round(y.shape[1]*nch/ms/1e6,1),'Gout/s'); break
PY
MRB_NO_DMMA=1 timeout 100 python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
from fractions import Fraction
import multirate_b200 as mr
rng=np.random.default_rng(0)
for ratio,nt,nch in ((Fraction(147,160),3528,4096),(Fraction(4,1),128,2048),(Fraction(1,1),63,4096)):
    h=rng.standard_normal(nt)
    x=torch.randn((nch,65536),device='cuda',dtype=torch.float64)
    f=mr.FIRFilter(h,ratio,nchannels=nch,sample_dtype=np.float64)
    for _ in range(3): f.filt(x)
    torch.cuda.synchronize(); f.set_timing(True)
    for _ in range(5): y=f.filt(x)
    torch.cuda.synchronize()
    ms=f.kernel_ms(); print('no dmma',ratio,nt,nch,f.last_kernel,round(ms,3),'ms',round(y.shape[1]*nch/ms/1e6,1),'Gout/s')
PY
