# ncu --set full of the dominant kernel of the other BASELINE configs (one launch each)
mkdir -p gpurun_out
for pair in "c1 k_stream" "c2 k_decim" "c3a k_unit" "c3b k_unit" "c4a k_table_fir" "c4f64 k_table_fir"; do
  set -- $pair
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/r1_$1 python bench.py --workload $1 --no-e2e --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
