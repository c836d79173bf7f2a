mkdir -p gpurun_out
out=gpurun_out/s4_sweep27.txt; : > $out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -6 >> $out
for w in c4a c4f; do
  echo "workload=$w" >> $out
  timeout 200 python bench.py --workload $w --no-e2e --no-cpu --steps 20 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['fp32']['frac'], d['roofline']['kernel_ms'], d['ms_per_step'])" >> $out
done
cat $out
