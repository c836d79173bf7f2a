mkdir -p gpurun_out
out=gpurun_out/s4_sweep21.txt; : > $out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -8 >> $out
cat $out
