mkdir -p gpurun_out
out=gpurun_out/s4_sweep8.txt; : > $out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -3 >> $out
for cfg in "3 4 0 1024" "3 4 8 1024" "3 4 0 1632" "3 4 0 816" "3 1 0 1024" "3 1 0 1632"; do
  set -- $cfg
  echo "variant=$1 tw=$2 dbg=$3 kt=$4" >> $out
  MRB_TILED_VARIANT=$1 MRB_TILED_TW=$2 MRB_TILED_DBG=$3 MRB_TILED_KT=$4 timeout 120 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['kernel'])" >> $out
done
cat $out
