o=gpurun_out; mkdir -p $o
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 380 -k "table_kernel_arbitrary_farrow and float64 and 0.918" > $o/r2_racecheck_dmma_full.txt 2>&1
grep -c "hazard" $o/r2_racecheck_dmma_full.txt; grep -E "Race reported|hazard|at .*\+0x|Write Thread|Read Thread|Current Value" $o/r2_racecheck_dmma_full.txt | head -60
