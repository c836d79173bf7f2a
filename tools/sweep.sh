mkdir -p gpurun_out
out=gpurun_out/s4_sweep12.txt; : > $out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -2 >> $out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/s4_bench_v9.json 2>>$out
cat gpurun_out/s4_bench_v9.json >> $out
cat $out
