o=gpurun_out; mkdir -p $o
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "rational or tiled or standard or c5 or shard or live_tap" 2>&1 | tail -5
timeout 200 python bench.py --only-main --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
