# round 2 final evidence: full GPU suite, default bench line, reference arm, launch lists + ncu captures of the new kernels
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -5 > $o/r2_pytest_gpu.txt; cat $o/r2_pytest_gpu.txt
timeout 1200 python bench.py --steps 20 --warmup 3 > $o/r2_bench_default.json 2> $o/r2_bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $o/r2_bench_reference_arm.json 2>> $o/r2_bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 8 --csv --log-file $o/r2_ncu_launches_c2.csv python tools/mma_one.py c2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_decim8 -s 4 -c 1 -f -o $o/r2_decim8 python tools/mma_one.py c2 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2_ncu_launches_c5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --only-main > /dev/null 2>&1
tail -3 $o/r2_bench_default.err | cut -c1-300
