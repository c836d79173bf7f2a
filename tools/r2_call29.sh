o=gpurun_out; mkdir -p $o
timeout 300 python tools/soak_new_kernels.py 1500 2>&1 | tail -6 | tee $o/r2_soak_new_kernels.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "host or pipeline or live_tap or table_kernel" 2>&1 | tail -4
for w in c4a c2 c4a64 c3b; do timeout 200 python bench.py --workload $w --only-main --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2))"; done
