o=gpurun_out; mkdir -p $o
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 450 -k "host_path_uses_fast_kernels or float64_integer_kinds or test_schedule_cache" 2>&1 | tail -4 > $o/r2_sanitizer_memcheck_tensor_paths.txt; cat $o/r2_sanitizer_memcheck_tensor_paths.txt
