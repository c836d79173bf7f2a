timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "stress or host_path or live_tap or long_stream or segment or golden or smoke or rational or c5 or shard or non_finite or tiled" 2>&1 | tail -6
for v in 1 0; do MRB_MMA_SPLIT=$v timeout 200 python bench.py --workload c5 --only-main --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 split=$v', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; done
for w in x160; do timeout 200 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))"; done
