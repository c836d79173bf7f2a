# Round-end measurement suite (one GPU): tests, headline bench, the other BASELINE configs, reference arm, ncu passes.
mkdir -p gpurun_out
o=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -3 > $o/r1_pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 3 > $o/r1_bench_c5_n1.json 2> $o/r1_bench_c5_n1.err
for w in c1 c2 c3a c3b c4a c4f c4a64 c4f64; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-e2e > $o/r1_bench_$w.json 2> $o/r1_bench_$w.err
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $o/r1_bench_reference_arm.json 2>> $o/r1_bench_c5_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r1_ncu_launches_c5.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tiled -s 2 -c 1 -f -o $o/r1_tiled_final python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
cat $o/r1_pytest_gpu.txt
for f in $o/r1_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(' value %.1f Msamples/s  ms/step %.3f  kernel %s frac %s fp32 %s e2e %s cpu %s' % (d['value'], d.get('ms_per_step',0), r.get('kernel'), r.get('frac'), (r.get('fp32') or {}).get('frac'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value')))
"; done
