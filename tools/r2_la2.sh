MRB_TRACE=1 timeout 100 python bench.py --workload c4a --only-main --steps 6 --warmup 3 --no-cpu --no-e2e 2>&1 | grep "look-ahead" | head -12
