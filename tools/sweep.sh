mkdir -p gpurun_out
out=gpurun_out/s4_sweep19.txt; : > $out
for wv in 6 3 12; do
for w in c3a c3b; do
  echo "workload=$w waves=$wv" >> $out
  MRB_UNIT_WAVES=$wv timeout 200 python bench.py --workload $w --no-e2e --no-cpu --steps 10 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['fp32']['frac'], d['ms_per_step'])" >> $out
done; done
python -c "import __graft_entry__ as g; g.smoke()" >> $out 2>&1
cat $out
