mkdir -p gpurun_out
out=gpurun_out/s4_sweep16.txt; : > $out
MRB_TRACE=1 timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -15 >> $out
for w in c4a c4f; do
  echo "workload=$w" >> $out
  timeout 200 python bench.py --workload $w --no-e2e --no-cpu --steps 10 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline'], d['ms_per_step'])" >> $out
done
cat $out
