for w in c3b c3a xr32; do MRB_MMA_PROF=1 timeout 100 python tools/mma_one.py $w 2>&1 | grep -E "mma prof|kernel=" | head -3 | cut -c1-700; done
