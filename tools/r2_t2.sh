timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 100 -k "one_shot or oneshot or golden or readme" 2>&1 | tail -4
timeout 200 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('c1_oneshot'))"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --only-main 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('c1_oneshot'), d['value'])"
