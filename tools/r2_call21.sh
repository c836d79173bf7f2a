o=gpurun_out; mkdir -p $o
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "table_kernel or full_size_c4 or host_pipeline or non_finite or arbitrary or farrow or stream" 2>&1 | tail -6
for w in c4a64 c4f64 x4ac; do timeout 100 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"; done
MRB_NO_SIDE_STREAM=1 timeout 100 python bench.py --workload c4a64 --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4a64 no side', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"
MRB_NO_SIDE_STREAM=1 timeout 100 python bench.py --workload x4ac --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('x4ac no side', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"
