o=gpurun_out; mkdir -p $o
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 10 --csv --log-file $o/r2_ncu_launches_c4a64.csv python tools/mma_one.py c4a64 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_table_fir -s 4 -c 1 -f -o $o/r2_table_f64_dmma python tools/mma_one.py c4a64 > /dev/null 2>&1
ls -la $o/r2_table_f64_dmma.ncu-rep
