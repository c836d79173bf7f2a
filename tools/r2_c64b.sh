timeout 600 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -6
for w in c5 x3ac; do timeout 200 python bench.py --workload $w --only-main --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['gpu_launches'])"; done
timeout 300 python tools/mma_dbg.py arb 2>&1 | tail -2
