timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "look_ahead or table_kernel or full_size_c4 or host_pipeline or arbitrary or farrow or mbarrier" 2>&1 | tail -5
for w in c4a c4f c4a64; do for v in 0 1; do if [ $v = 1 ]; then export MRB_NO_LOOKAHEAD=1; else unset MRB_NO_LOOKAHEAD; fi; timeout 100 python bench.py --workload $w --only-main --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w no_lookahead=$v', round(d['value'],1), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), round(d['ms_per_step'],4))"; done; done
unset MRB_NO_LOOKAHEAD
timeout 300 python tools/mma_dbg.py arb 2>&1 | tail -2
