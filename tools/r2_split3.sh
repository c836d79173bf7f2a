o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -5
timeout 400 python tools/soak_new_kernels.py 1000 2>&1 | tail -7 | tee $o/r2_soak_new_kernels.txt
timeout 300 python tools/mma_dbg.py arb 2>&1 | tail -1
timeout 200 python bench.py --only-main --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4), d['clocks'])"
