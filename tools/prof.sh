mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decim -s 2 -c 1 -f -o gpurun_out/s4_decim2 python bench.py --workload c2 --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/s4_prof.log 2>&1
tail -1 gpurun_out/s4_prof.log | cut -c1-100
