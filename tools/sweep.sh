mkdir -p gpurun_out
out=gpurun_out/s4_sweep24.txt; : > $out
for cfg in "64 2" "64 3" "32 3" "128 3" "32 4" "16 4" "256 2"; do
  set -- $cfg
  echo "block_mib=$1 streams=$2" >> $out
  MRB_HOST_BLOCK_MIB=$1 MRB_HOST_STREAMS=$2 timeout 200 python bench.py --no-cpu --steps 3 --warmup 3 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['e2e']['ms_per_step'])" >> $out
done
cat $out
