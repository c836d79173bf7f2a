o=gpurun_out; mkdir -p $o
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -s -k "live_tap" 2>&1 | tail -12 > $o/r2_pytest_gpu_12a.txt
cat $o/r2_pytest_gpu_12a.txt
timeout 300 python -m pytest tests -m gpu -q --timeout 100 -x 2>&1 | tail -6 > $o/r2_pytest_gpu_12.txt
cat $o/r2_pytest_gpu_12.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r2_bench_default4.json 2> $o/r2_bench_default4.err
tail -3 $o/r2_bench_default4.err
