"""Rewrites the per-config rows of profiles/README.md's round-2 table from profiles/r2_bench_default.json."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_default.json")).read().strip().splitlines()[-1])
c = d["configs"]
def row(k):
    v = c[k]
    return v["value"] / 1e3, v["kernel_ms"], v.get("frac"), v["hbm_frac"], v["e2e"]["value"] / 1e3
p = os.path.join(ROOT, "profiles", "README.md")
s = open(p).read()
a = s.index("| **C5** rational 147//160 c64, 8192 ch (headline)")
b = s.index("Earlier runs of the same command on other boxes are kept")
rows = []
rows.append("| **C5** rational 147//160 c64, 8192 ch (headline) | **`k_mma_fir<32, split>`** (`%s`) | **%.1f** | %.3f | HBM **%.3f** | %.3f | %.2f | 308.4 (`k_tiled_c64`, now the CUDA-core arm: 311) |" % (d["roofline"]["kernel"], d["value"] / 1e3, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["frac"], d["e2e"]["value"] / 1e3))
v = row("c1"); rows.append("| C1 README one-shot shape, 1 ch f32, 1e6 | `k_stream` (1 channel is below the tensor-core kernel's 48-channel floor) | %.1f | %.3f | latency | — | %.1f | 24.0 |" % (v[0], v[1], v[4]))
v = row("c2"); rows.append("| C2 decimator 1//8 × 256, 1024 ch c64 | **`k_decim8`** (lane per channel, launch-constant taps) | **%.1f** | %.3f | FP32 **%.3f** | %.2f | %.2f (H2D alone: 8 input bytes per output byte) | 34.3 |" % (v[0], v[1], v[2], v[3], v[4]))
v = row("c3a"); rows.append("| C3a interpolator 4//1 × 128, 4096 ch f32 | **`k_mma_fir` (resident tile)** | **%.1f** | %.3f | HBM **%.2f** | %.2f | %.1f | 857.9 |" % (v[0], v[1], v[2], v[3], v[4]))
v = row("c3b"); rows.append("| C3b standard × 128, 4096 ch f32 | **`k_mma_fir` (resident tile)** | **%.1f** | %.3f | FP32 **%.2f** (tensor cores) | %.2f | %.1f | 224.2 |" % (v[0], v[1], v[2], v[3], v[4]))
v = row("c4a"); rows.append("| C4 arbitrary 0.918734, 1024 ch f32 | **`k_mma_fir`** | **%.1f** | %.3f | FP32 %.2f | %.2f | %.1f | 117.3 |" % (v[0], v[1], v[2], v[3], v[4]))
v = row("c4f"); rows.append("| C4 farrow, 1024 ch f32 | **`k_mma_fir`** | **%.1f** | %.3f | FP32 %.2f | %.2f | %.1f | 118.3 |" % (v[0], v[1], v[2], v[3], v[4]))
va, vf = row("c4a64"), row("c4f64"); rows.append("| C4 arbitrary / farrow, 1024 ch f64 | **`k_table_fir<f64>` on `mma.sync.m8n8k4.f64`** | **%.1f / %.1f** | %.3f / %.3f | FP64 **%.2f / %.2f** | %.2f | %.1f / %.1f | 58.1 / 58.0 |" % (va[0], vf[0], va[1], vf[1], va[2], vf[2], va[3], va[4], vf[4]))
v = row("xr32"); rows.append("| extra: rational 147//160 **f32**, 8192 ch (the README dtype) | **`k_mma_fir`** (147 L2-resident tiles) | **%.1f** | %.3f | HBM **%.3f** | %.3f | %.1f | 58 (`k_stream`) |" % (v[0], v[1], v[2], v[3], v[4]))
v = row("x4a8k"); rows.append("| extra: arbitrary f32, **8192** ch | **`k_mma_fir`** | **%.1f** | %.3f | FP32 %.2f | %.2f | %.1f | — |" % (v[0], v[1], v[2], v[3], v[4]))
v = row("xr64"); rows.append("| extra: rational 147//160 **f64**, 4096 ch | **`k_table_fir<f64>` DMMA, closed-form schedule** (`int_f64_dmma`) | **%.1f** | %.3f | HBM **%.2f** | %.2f | %.1f | 31 (`k_stream`) |" % (v[0], v[1], v[2], v[3], v[4]))
st = d["stream_2e31"]; rows.append("| 2^31-sample c64 stream, 147//160, one GPU (`stream_2e31`) | `LongStream` (plan and handles built once) | %.1f (%.1f ms) | — | — | %.2f | — | 156.7 |" % (st["value"] / 1e3, st["ms"], st["algorithmic_gbs_per_gpu"] / 6548.5))
o = d["c1_oneshot"]; rows.append("| C1 one-shot end to end (`c1_oneshot`: `filt(h, x, 147//160)` on numpy, H2D + D2H included) | | repeated call (kept handle, reset): **%.2f ms** median; new handle every call: %.1f ms; first call of the process %.1f ms (README: 56.9 ms on 2014 hardware) | | | | | — |" % (o["seconds_median"] * 1e3, o.get("seconds_median_new_handle_every_call", float("nan")) * 1e3, o["first_call_seconds"] * 1e3))
s = s[:a] + "\n".join(rows) + "\n\n" + s[b:]
open(p, "w").write(s)
print("\n".join(rows))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline_1t"]["value"])
