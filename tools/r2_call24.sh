o=gpurun_out; mkdir -p $o
timeout 60 python tools/mma_one.py c2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_decim8 -s 4 -c 1 -f -o $o/r2_decim8 python tools/mma_one.py c2 > /dev/null 2>&1
ls -la $o/r2_decim8.ncu-rep
