o=gpurun_out; mkdir -p $o
nvidia-smi -L | wc -l
for n in 8 4 2; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --only-main > $o/r2_bench_n$n.json 2> $o/r2_bench_n$n.err; echo "n$n rc=$?"
python -c "
import json
d=json.loads(open('$o/r2_bench_n$n.json').read().strip().splitlines()[-1]); print($n, round(d['value'],1), d['unit'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'], d['clocks']['sm_mhz'])"
done
