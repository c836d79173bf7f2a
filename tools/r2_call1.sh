# round 2, GPU call 1: parity suite (incl. the host-pipeline race regression), tcgen05 probe, copy ceiling, sanitizer
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15 > $o/r2_pytest_gpu_1.txt
timeout 120 tools/mma_probe > $o/r2_mma_probe.txt 2>&1
timeout 300 tools/h2d_ceiling 1024 4 0 > $o/r2_h2d_ceiling_n1.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "host_pipeline_multi_slice and float32 and 80000" 2>&1 | tail -25 > $o/r2_sanitizer_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "host_pipeline_multi_slice and 80000" 2>&1 | tail -25 > $o/r2_sanitizer_memcheck.txt
cat $o/r2_pytest_gpu_1.txt $o/r2_mma_probe.txt $o/r2_h2d_ceiling_n1.txt
tail -8 $o/r2_sanitizer_racecheck.txt $o/r2_sanitizer_memcheck.txt
