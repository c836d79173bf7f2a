# round 2 final evidence: full GPU suite, default bench line, reference arm, launch lists
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -5 > $o/r2_pytest_gpu.txt; cat $o/r2_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --steps 20 --warmup 3 > $o/r2_bench_default.json 2> $o/r2_bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $o/r2_bench_reference_arm.json 2>> $o/r2_bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 8 --csv --log-file $o/r2_ncu_launches_c2.csv python tools/mma_one.py c2 > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2_ncu_launches_c5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --only-main > /dev/null 2>&1
tail -3 $o/r2_bench_default.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('c5', round(d['value'],1), d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), v['kernel'], v['kernel_ms'], v.get('frac'), 'e2e', round(v['e2e']['value'],1) if v.get('e2e') else None, v['clocks']['sm_mhz'], v['clocks']['reasons'])
print(d['c1_oneshot']['seconds_median'], d['stream_2e31']['value'])
PY
