# round 2 evidence: bench line, launch lists, ncu full captures of the dominant kernels
o=gpurun_out; mkdir -p $o
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r2_bench_default.json 2> $o/r2_bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $o/r2_bench_reference_arm.json 2>> $o/r2_bench_default.err
for w in c4a c3b xr32; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file $o/r2_ncu_launches_$w.csv python tools/mma_one.py $w > /dev/null 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mma_fir -s 4 -c 1 -f -o $o/r2_ncu_mma_$w python tools/mma_one.py $w > /dev/null 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2_ncu_launches_c5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --only-main > /dev/null 2>&1
tail -3 $o/r2_bench_default.err | cut -c1-300
