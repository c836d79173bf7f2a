o=gpurun_out; mkdir -p $o
timeout 100 python tools/mma_one.py c5
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_mma_fir_c5 python tools/mma_one.py c5 > /dev/null 2>&1
ls -la $o/r2_mma_fir_c5.ncu-rep
MRB_MMA_PROF=1 timeout 100 python tools/mma_one.py c5 2>&1 | grep "mma prof" | head -1 | cut -c1-460
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_mma_fir_c3b python tools/mma_one.py c3b > /dev/null 2>&1
