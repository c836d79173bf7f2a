mkdir -p gpurun_out
out=gpurun_out/s4_sweep18.txt; : > $out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -3 >> $out
for ob in 3 1 2; do
  echo "ob=$ob" >> $out
  MRB_TILED_OB=$ob timeout 120 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['kernel'])" >> $out
done
cat $out
