o=gpurun_out; mkdir -p $o
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_mma_fir_v2 python tools/mma_check2.py > /dev/null 2>&1
ls -la $o/r2_mma_fir_v2.ncu-rep
