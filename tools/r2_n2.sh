o=gpurun_out; mkdir -p $o
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $o/r2_bench_n2_full.json 2> $o/r2_bench_n2_full.err; echo "n2 rc=$?"
tail -2 $o/r2_bench_n2_full.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_full.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), d['roofline']['kernel'], 'e2e', round(d['e2e']['value'],1))
for k,v in d['configs'].items(): print(k, round(v['value'],1) if 'value' in v else v)
print(d['stream_2e31'].get('value'), d['stream_2e31'].get('segments'), d.get('c1_oneshot',{}).get('seconds_median'))
PY
