mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_table_fir -s 2 -c 1 -f -o gpurun_out/s4_table python bench.py --workload c4a --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/s4_prof.log 2>&1
tail -2 gpurun_out/s4_prof.log | cut -c1-200
