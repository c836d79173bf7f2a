o=gpurun_out; mkdir -p $o
for w in c3b c4a xr32 c3a; do timeout 100 python tools/mma_one.py $w 2>&1 | tail -1; done > $o/r2_two_issuers.txt
timeout 100 python tools/mma_one.py c4a 8192 2>&1 | tail -1 >> $o/r2_two_issuers.txt
MRB_MMA_PROF=1 timeout 100 python tools/mma_one.py c3b 2>&1 | tail -2 >> $o/r2_two_issuers.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -8 > $o/r2_pytest_gpu_10.txt
cat $o/r2_two_issuers.txt $o/r2_pytest_gpu_10.txt
