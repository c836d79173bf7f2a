o=gpurun_out; mkdir -p $o
for w in c5 x3bc x3ac x4ac x160; do for v in 1 0; do MRB_MMA_C64=$v timeout 200 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w mma_c64=$v', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))"; done; done
timeout 600 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -15
