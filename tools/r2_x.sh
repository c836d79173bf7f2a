for w in x3ac x3bc; do for v in 0 1; do if [ $v = 1 ]; then export MRB_C64_UNIT_FIRST=1; else unset MRB_C64_UNIT_FIRST; fi; timeout 200 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w unit_first=$v', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"; done; done
unset MRB_C64_UNIT_FIRST
timeout 100 python - <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0,'.')
from fractions import Fraction
import multirate_b200 as mr
rng=np.random.default_rng(0)
for nt in (32, 48, 64):
    h=rng.standard_normal(nt).astype(np.float32)
    x=torch.randn((4096,65536),device='cuda',dtype=torch.complex64)
    for pol in (0,2):
        f=mr.FIRFilter(h,Fraction(1,1),nchannels=4096,sample_dtype=np.complex64)
        f.set_kernel_policy(pol)
        for _ in range(3): f.filt(x)
        torch.cuda.synchronize(); f.set_timing(True)
        for _ in range(5): y=f.filt(x)
        torch.cuda.synchronize()
        ms=f.kernel_ms(); print('standard c64',nt,'taps policy',pol,f.last_kernel,round(ms,3),'ms',round(y.shape[1]*4096/ms/1e6,1),'Gout/s')
PY
