# The "extra" workloads of bench.py (same kernels on the other sample type / mirrored ratio), one GPU, device-timed only.
mkdir -p gpurun_out
for w in x160 x2f x3ac x3bc x4ac x4fc; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1_bench_$w.json 2> gpurun_out/r1_bench_$w.err
  python -c "
import json
d=json.loads(open('gpurun_out/r1_bench_$w.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$w %.1f Msamples/s  kernel %s %.4f ms  hbm %.3f  fp32 %.3f (%.1f TFLOP/s)' % (d['value'], r['kernel'], r['kernel_ms'], r['frac'], r['fp32']['frac'], r['fp32']['achieved_tflops']))
"
done
