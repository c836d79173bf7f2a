o=gpurun_out; mkdir -p $o
timeout 60 python tools/mma_one.py c2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_decim8 -s 4 -c 1 -f -o $o/r2_decim8 python tools/mma_one.py c2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_table_fir -s 4 -c 1 -f -o $o/r2_table_f64_dmma3 python tools/mma_one.py c4a64 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 10 --csv --log-file $o/r2_ncu_launches_c4a64.csv python tools/mma_one.py c4a64 > /dev/null 2>&1
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 280 -k "table_kernel_arbitrary_farrow and float64" 2>&1 | tail -4 > $o/r2_sanitizer_racecheck_dmma.txt; cat $o/r2_sanitizer_racecheck_dmma.txt
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 280 -k "(table_kernel_arbitrary_farrow and float64) or decimator_m8_lane" 2>&1 | tail -4 > $o/r2_sanitizer_memcheck_new_kernels.txt; cat $o/r2_sanitizer_memcheck_new_kernels.txt
ls -la $o/*.ncu-rep | tail -3
