o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -5 > $o/r2_pytest_gpu.txt; cat $o/r2_pytest_gpu.txt
