o=gpurun_out; mkdir -p $o
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "non_finite" 2>&1 | tail -8 > $o/r2_pytest_nonfinite.txt; cat $o/r2_pytest_nonfinite.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $o/r2_bench_n2.json 2> $o/r2_bench_n2.err; echo "n2 rc=$?"
tail -3 $o/r2_bench_n2.err | cut -c1-300
timeout 300 tools/h2d_ceiling 1024 4 0 > $o/r2_h2d_ceiling_n2.txt 2>&1; cat $o/r2_h2d_ceiling_n2.txt
