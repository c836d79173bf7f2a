mkdir -p gpurun_out
out=gpurun_out/s4_sweep15.txt; : > $out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -4 >> $out
for w in c2; do
  echo "workload=$w" >> $out
  timeout 200 python bench.py --workload $w --no-e2e --no-cpu --steps 20 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline'], d['ms_per_step'])" >> $out
done
cat $out
