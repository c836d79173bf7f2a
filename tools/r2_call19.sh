o=gpurun_out; mkdir -p $o
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "table_kernel or full_size_c4 or host_pipeline_multi or non_finite or arbitrary_and_farrow" 2>&1 | tail -6
for v in 8 4; do for w in c4a64 c4f64; do MRB_DMMA_WARPS=$v timeout 100 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w warps $v', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"; done; done
MRB_NO_DMMA=1 timeout 100 python bench.py --workload c4a64 --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4a64 dfma', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"
