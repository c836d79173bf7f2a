o=gpurun_out; mkdir -p $o
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "tensor_core or mma or table_kernel or unit or standard or interpolator or rational or host" 2>&1 | tail -4
for w in c3b c3a xr32 c4a; do timeout 100 python tools/mma_one.py $w 2>&1 | tail -1; MRB_MMA_FWD=0 timeout 100 python tools/mma_one.py $w 2>&1 | tail -1 | sed 's/^/  (fwd off) /'; done
timeout 100 python tools/mma_one.py c4a 8192 2>&1 | tail -1
timeout 400 python tools/mma_dbg.py arb 2>&1 | tail -3
