o=gpurun_out; mkdir -p $o
timeout 60 python tools/mma_one.py c4a64
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_table_fir -s 4 -c 1 -f -o $o/r2_table_f64_2ch python tools/mma_one.py c4a64 > /dev/null 2>&1
ls -la $o/r2_table_f64_2ch.ncu-rep
