o=gpurun_out; mkdir -p $o
timeout 60 tools/dfma_probe 2>&1 | tee $o/r2_dfma_probe.txt
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 2>&1 | tail -4
