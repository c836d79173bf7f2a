o=gpurun_out; mkdir -p $o
for w in c3b c4a xr32 c3a; do timeout 100 python tools/mma_one.py $w 2>&1 | tail -1; done > $o/r2_waiter.txt
timeout 100 python tools/mma_one.py c4a 8192 2>&1 | tail -1 >> $o/r2_waiter.txt
cat $o/r2_waiter.txt
timeout 250 python tools/mma_dbg.py arb 2>&1 | grep -v CUDAEvent | tail -2 | cut -c1-300
timeout 250 python tools/mma_dbg.py 4 2>&1 | grep -v CUDAEvent | tail -2 | cut -c1-300
timeout 300 python -m pytest tests -m gpu -q --timeout 100 -x 2>&1 | tail -6 > $o/r2_pytest_gpu_15.txt
cat $o/r2_pytest_gpu_15.txt
