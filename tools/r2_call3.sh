o=gpurun_out; mkdir -p $o
MRB_MMA_PROF=1 timeout 200 python tools/mma_check2.py > $o/r2_mma_prof.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/r2_ncu_launches_c4a.csv python tools/mma_check2.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_mma_fir python tools/mma_check2.py > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15 > $o/r2_pytest_gpu_3.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r2_bench_default.json 2> $o/r2_bench_default.err
cat $o/r2_mma_prof.txt; grep -E "k_mma|k_generic|k_hist|k_table" $o/r2_ncu_launches_c4a.csv | cut -d, -f5,15- | head -30
cat $o/r2_pytest_gpu_3.txt; tail -5 $o/r2_bench_default.err; head -c 6000 $o/r2_bench_default.json
