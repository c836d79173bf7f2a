mkdir -p gpurun_out
out=gpurun_out/s4_sweep25.txt; : > $out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "decimator or stress or host_path" 2>&1 | tail -3 >> $out
for wv in 4; do
  echo "waves=$wv" >> $out
  MRB_DEC_WAVES=$wv timeout 200 python bench.py --workload c2 --no-e2e --no-cpu --steps 20 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['fp32']['frac'], d['ms_per_step'])" >> $out
done
cat $out
