o=gpurun_out; mkdir -p $o
MRB_MMA_PROF=1 timeout 200 python tools/mma_check2.py > $o/r2_mma_prof3.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/r2_ncu_launches_c4a_v3.csv python tools/mma_check2.py > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > $o/r2_pytest_gpu_6.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r2_bench_default3.json 2> $o/r2_bench_default3.err
cat $o/r2_mma_prof3.txt; grep -E "k_mma|k_generic|k_hist|k_head" $o/r2_ncu_launches_c4a_v3.csv | cut -d, -f5,15- | head -8
cat $o/r2_pytest_gpu_6.txt; tail -5 $o/r2_bench_default3.err
