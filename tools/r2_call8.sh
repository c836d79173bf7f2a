o=gpurun_out; mkdir -p $o
MRB_TRACE=1 timeout 300 python tools/mma_check3.py > $o/r2_mma_check_int.txt 2>&1
timeout 200 python tools/mma_check2.py > $o/r2_mma_check2_v5.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 > $o/r2_pytest_gpu_8.txt
cat $o/r2_mma_check_int.txt | tail -30; cat $o/r2_mma_check2_v5.txt; cat $o/r2_pytest_gpu_8.txt
