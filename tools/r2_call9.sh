o=gpurun_out; mkdir -p $o
for w in c3b c4a; do MRB_MMA_PROF=1 timeout 100 python tools/mma_one.py $w 2>&1 | tail -2 > $o/r2_prof_$w.txt; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_ncu_mma_c4a python tools/mma_one.py c4a > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_ncu_mma_c3b python tools/mma_one.py c3b > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_ncu_mma_c4a8k python tools/mma_one.py c4a 8192 > /dev/null 2>&1
cat $o/r2_prof_c3b.txt $o/r2_prof_c4a.txt; ls -la $o/*.ncu-rep
