o=gpurun_out; mkdir -p $o
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "decim or live_tap or streaming or host" 2>&1 | tail -8
for w in c2; do timeout 100 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"; done
MRB_DECIM8=0 timeout 100 python bench.py --workload c2 --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 old', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4))"
