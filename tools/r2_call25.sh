o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -x --timeout 200 2>&1 | tail -4
for w in c2; do timeout 100 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), d['clocks'])"; done
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 280 -k "decimator_m8_lane" 2>&1 | tail -5
