o=gpurun_out; mkdir -p $o
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "arbitrary_and_farrow or tensor_core or table_kernel" 2>&1 | tail -6 > $o/r2_pytest_gpu_11a.txt
cat $o/r2_pytest_gpu_11a.txt
timeout 300 python -m pytest tests -m gpu -q --timeout 100 -x 2>&1 | tail -8 > $o/r2_pytest_gpu_11.txt
cat $o/r2_pytest_gpu_11.txt
