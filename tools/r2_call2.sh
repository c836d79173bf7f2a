o=gpurun_out; mkdir -p $o
timeout 120 tools/mma_probe > $o/r2_mma_probe2.txt 2>&1
MRB_TRACE=1 timeout 300 python tools/mma_check.py > $o/r2_mma_check.txt 2>&1
echo "mma_check rc=$?" >> $o/r2_mma_check.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15 > $o/r2_pytest_gpu_2.txt
tail -16 $o/r2_mma_probe2.txt; cat $o/r2_mma_check.txt; cat $o/r2_pytest_gpu_2.txt
