mkdir -p gpurun_out
out=gpurun_out/s4_sweep22.txt; : > $out
timeout 800 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -12 >> $out
timeout 120 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])" >> $out
cat $out
