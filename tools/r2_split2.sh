for w in xr32 c3b c3a c4a; do timeout 100 python tools/mma_one.py $w 2>&1 | tail -1; done
MRB_MMA_SPLIT=0 timeout 100 python tools/mma_one.py c5 2>&1 | tail -1
timeout 100 python tools/mma_one.py c5 2>&1 | tail -1
