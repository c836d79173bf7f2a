#!/usr/bin/env python
"""bench.py -- streaming polyphase FIR throughput on B200 (the driver's contract, tier (4) reading).

A "step" is one pass of the hot path over one batch of synthetic input: one 65,536-sample chunk per channel
through a stateful FIRFilter (history, phase and deficit carried on the device between steps).

Headline workload (N=1 and every N, weak scaling): BASELINE.json configs[4]'s per-GPU shard --
FIRRational 147//160, 3528-tap Kaiser low-pass (Float32 taps), 8192 channels of Complex64 per GPU
(65,536 channels over 8 GPUs), 64K-sample chunks: `value`, `roofline`, `e2e`, `cpu_baseline` are quoted on it.

The same JSON line also carries, under `configs`, every other BASELINE config measured in the same run (device-timed
value, dominant kernel and its time, roofline fraction against the bound SURVEY 8d names, end-to-end value, clocks):
c1 (README one-shot, plus its host-to-host one-shot time beside the README's 0.0569 s), c2, c3a, c3b, c4a, c4f, c4a64,
c4f64, and `stream` = the 2^31-sample stream of configs[4] split into M-aligned segments over the N ranks
(segment_bounds + seek semantics, tap-length halo, no collective).  `--workload X` makes X the headline instead;
`--only-main` skips the extra configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 1 << 16

# name -> (description, ratio, ntaps, cutoff, beta, gain, sample dtype, channels per GPU, Nphi, polyorder)
WORKLOADS = {
    "c5": ("FIRRational 147//160, 3528 taps, 8192 ch/GPU complex64, 64K-sample chunks (BASELINE configs[4] per-GPU shard)",
           Fraction(147, 160), 3528, 0.5 / 147, 7.8562, 1.0, np.complex64, 8192, None, None),
    "c1": ("README benchmark: FIRRational 147//160, 3528 taps, 1 ch float32, 1e6 samples (BASELINE configs[0])",
           Fraction(147, 160), 3528, 0.5 / 147, 7.8562, 1.0, np.float32, 1, None, None),
    "c2": ("FIRDecimator 1//8, 256 taps, 1024 ch complex64, 64K-sample chunks (BASELINE configs[1])",
           Fraction(1, 8), 256, 0.5 / 8, 7.8562, 1.0, np.complex64, 1024, None, None),
    "c3a": ("FIRInterpolator 4//1, 128 taps, 4096 ch float32 (BASELINE configs[2])",
            Fraction(4, 1), 128, 0.5 / 4, 7.8562, 4.0, np.float32, 4096, None, None),
    "c3b": ("FIRStandard, 128 taps, 4096 ch float32 (BASELINE configs[2])",
            Fraction(1, 1), 128, 0.25, 7.8562, 1.0, np.float32, 4096, None, None),
    "c4a": ("FIRArbitrary rate 0.918734, Nphi 32, 2336 taps, 1024 ch float32 (BASELINE configs[3])",
            0.918734, 2336, 0.45 / 32, 5.6533, 32.0, np.float32, 1024, 32, None),
    "c4f": ("FIRFarrow rate 0.918734, Nphi 32, 2336 taps, order 4, 1024 ch float32 (BASELINE configs[3])",
            0.918734, 2336, 0.45 / 32, 5.6533, 32.0, np.float32, 1024, 32, 4),
    "c4a64": ("FIRArbitrary rate 0.918734, Nphi 32, 2336 taps, 1024 ch float64 (BASELINE configs[3])",
              0.918734, 2336, 0.45 / 32, 5.6533, 32.0, np.float64, 1024, 32, None),
    "c4f64": ("FIRFarrow rate 0.918734, Nphi 32, 2336 taps, order 4, 1024 ch float64 (BASELINE configs[3])",
              0.918734, 2336, 0.45 / 32, 5.6533, 32.0, np.float64, 1024, 32, 4),
    # extras (not BASELINE configs): the same kernels on the other sample type / the mirrored ratio / more channels
    "x160": ("extra: FIRRational 160//147, 3840 taps, 8192 ch complex64", Fraction(160, 147), 3840, 0.5 / 160, 7.8562, 1.0,
             np.complex64, 8192, None, None),
    "x2f": ("extra: FIRDecimator 1//8, 256 taps, 4096 ch float32", Fraction(1, 8), 256, 0.5 / 8, 7.8562, 1.0,
            np.float32, 4096, None, None),
    "x3ac": ("extra: FIRInterpolator 4//1, 128 taps, 4096 ch complex64", Fraction(4, 1), 128, 0.5 / 4, 7.8562, 4.0,
             np.complex64, 4096, None, None),
    "x3bc": ("extra: FIRStandard, 128 taps, 4096 ch complex64", Fraction(1, 1), 128, 0.25, 7.8562, 1.0,
             np.complex64, 4096, None, None),
    "x4ac": ("extra: FIRArbitrary rate 0.918734, Nphi 32, 2336 taps, 1024 ch complex64", 0.918734, 2336, 0.45 / 32,
             5.6533, 32.0, np.complex64, 1024, 32, None),
    "x4fc": ("extra: FIRFarrow rate 0.918734, Nphi 32, 2336 taps, order 4, 1024 ch complex64", 0.918734, 2336, 0.45 / 32,
             5.6533, 32.0, np.complex64, 1024, 32, 4),
    "x4a8k": ("extra: FIRArbitrary rate 0.918734, Nphi 32, 2336 taps, 8192 ch float32", 0.918734, 2336, 0.45 / 32,
              5.6533, 32.0, np.float32, 8192, 32, None),
    "xr32": ("extra: FIRRational 147//160, 3528 taps, 8192 ch float32 (the README dtype, multichannel)", Fraction(147, 160), 3528,
             0.5 / 147, 7.8562, 1.0, np.float32, 8192, None, None),
    "xr64": ("extra: FIRRational 147//160, 3528 taps, 4096 ch float64 (integer kinds on the FP64 tensor-core kernel)", Fraction(147, 160),
             3528, 0.5 / 147, 7.8562, 1.0, np.float64, 4096, None, None),
}
EXTRA_CONFIGS = ["c1", "c2", "c3a", "c3b", "c4a", "c4f", "c4a64", "c4f64", "xr32", "x4a8k", "xr64"]   # the BASELINE configs + three extras
# taps per output and the bound SURVEY 8d assigns (roofline denominators: measured HBM; nominal FP32 / FP64 FMA rate)
TAPS_PER_OUT = {"c5": 24, "c1": 24, "c2": 256, "c3a": 32, "c3b": 128, "c4a": 73, "c4f": 73, "c4a64": 73, "c4f64": 73,
                "x160": 24, "x2f": 256, "x3ac": 32, "x3bc": 128, "x4ac": 73, "x4fc": 73, "x4a8k": 73, "xr32": 24, "xr64": 24}
# (c3b / c4a / c4f were FP32-FMA bound on CUDA cores per SURVEY 8d; on the tensor-core kernel their roof is HBM: both are given)
BOUND = {"c5": "hbm", "c1": "latency", "c2": "fp32", "c3a": "hbm", "c3b": "fp32", "c4a": "fp32", "c4f": "fp32",
         "c4a64": "fp64", "c4f64": "fp64", "x160": "hbm", "x2f": "fp32", "x3ac": "hbm", "x3bc": "fp32", "x4ac": "fp32",
         "x4fc": "fp32", "x4a8k": "fp32", "xr32": "hbm", "xr64": "hbm"}
FP32_PEAK_TF = 148 * 128 * 2 * 1.965e9 / 1e12          # nominal CUDA-core FP32, TFLOP/s
README_ONESHOT_S = 0.056938961                          # README.md:190-193, the only number upstream publishes


def design_taps(ntaps, cutoff, beta, gain):
    """firdes(numtaps, cutoff, kaiser, beta) twin (reference src/FIRDesign.jl:52,76-86); taps are an input."""
    M = ntaps - 1
    n = np.arange(ntaps, dtype=np.float64)
    return (2 * cutoff * np.sinc(2 * cutoff * (n - M / 2)) * np.kaiser(ntaps, beta) * gain).astype(np.float32)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms from BEFORE the warm-up until the end of the run; a
    timed region asks for the samples that fall inside its wall-clock window (B200_PROFILING.md clocks line).  Regions
    shorter than a few sampling periods are followed by an untimed continuation of the same step loop so that the
    window holds samples taken under that load (`window` says which)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def _rows(self):
        import datetime
        rows = []
        if self.proc is None:
            return rows
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), [v.lower().startswith("active") for v in f[5:9]]))
            except ValueError:
                continue
        return rows

    def window(self, t0, t1, label="timed region"):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": label}
        rows = [r for r in self._rows() if t0 - 0.03 <= r[0] <= t1 + 0.03]
        if rows:
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            reasons = sorted({n for r in rows for n, a in zip(names, r[3]) if a})
            out.update(sm_mhz=float(np.median([r[1] for r in rows])), sm_max_mhz=max(r[2] for r in rows), reasons=reasons,
                       samples=len(rows))
        return out

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            os.unlink(self.path)
        except OSError:
            pass
        self.proc = None


def make_filter(mr, w, nch=None, device=0):
    desc, ratio, ntaps, cutoff, beta, gain, tx, nch_default, nphi, po = WORKLOADS[w]
    h = design_taps(ntaps, cutoff, beta, gain)
    if np.dtype(tx) == np.float64:
        h = h.astype(np.float64)                          # Float64 taps with Float64 samples
    nch = nch or nch_default
    if isinstance(ratio, float):
        return mr.FIRFilter(h, ratio, nphi, po, nchannels=nch, sample_dtype=tx, device=device), h
    return mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx, device=device), h


def chunk_len(w):
    return 1_000_000 if w == "c1" else CHUNK


# ------------------------------------------------------------------------------------------------
# CPU baseline: the C restatement of the reference loops (oracle/mr_oracle.c).
# ------------------------------------------------------------------------------------------------
def cpu_port_run(w, steps, warmup, target_s=6.0, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as co
    try:
        co.load(native=True)
        native = True
    except Exception:
        native = False
    desc, ratio, ntaps, cutoff, beta, gain, tx, nch_default, nphi, po = WORKLOADS[w]
    h = design_taps(ntaps, cutoff, beta, gain)
    if np.dtype(tx) == np.float64:
        h = h.astype(np.float64)
    threads = threads or os.cpu_count() or 1
    n = chunk_len(w)
    rng = np.random.default_rng(0x4D520000)

    def mk(nch):
        if isinstance(ratio, float):
            pn = None
            if po is not None:
                import multirate_oracle as mo
                pn = mo.pfb2pnfb(mo.taps2pfb(h, nphi), po)
            return co.COracleFilter("farrow" if po is not None else "arbitrary", h, tx, nch, rate=ratio, Nphi=nphi,
                                    polyorder=po or 0, pnfb=pn, native=native)
        L, M = ratio.numerator, ratio.denominator
        kind = "standard" if ratio == 1 else "decimator" if L == 1 else "interpolator" if M == 1 else "rational"
        return co.COracleFilter(kind, h, tx, nch, L, M, native=native)

    def data(nch):
        x = rng.random((nch, n), dtype=np.float32)
        if np.dtype(tx).kind == "c":
            x = (x + 1j * rng.random((nch, n), dtype=np.float32)).astype(np.complex64)
        return x.astype(tx)

    # calibrate the sample so that one step is ~target_s/steps of wall time
    nch = min(threads, nch_default)
    f, x = mk(nch), data(nch)
    t0 = time.perf_counter(); y = f.filt(x, threads); dt = time.perf_counter() - t0
    rounds = max(1, int(target_s / max(steps, 1) / max(dt, 1e-6)))      # dt = one channel per thread
    want = min(nch_default, rounds * nch) if nch_default > nch else nch
    f, x = mk(want), data(want)
    out = np.empty((want, y.shape[1] + 8), dtype=y.dtype)
    for _ in range(warmup):
        f.filt(x, threads, out=out)
    t0 = time.perf_counter()
    total = 0
    for _ in range(steps):
        total += f.filt(x, threads, out=out).shape[1] * want
    dt = time.perf_counter() - t0
    # OpenMP runs over channels: a workload with fewer channels than host threads uses that many threads
    return {"value": total / dt / 1e6, "unit": "Msamples/s", "cores": min(threads, want), "kind": "port",
            "sample": "%d channels x %d samples per step, %d steps, C restatement of the reference loops "
                      "(oracle/mr_oracle.c, gcc -O3 %s, OpenMP over channels, %d thread(s)); Julia reference not runnable: no julia in image"
                      % (want, n, steps, "-march=native" if native else "-march=x86-64-v3", min(threads, want)),
            "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = args.workload
    r = cpu_port_run(w, args.steps, max(args.warmup, 1), target_s=20.0)
    desc, ratio, ntaps, cutoff, beta, gain, tx, nch_default, nphi, po = WORKLOADS[w]
    line = {"metric": "output Msamples/s (multichannel, device-timed)", "value": r["value"], "unit": "Msamples/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(w), "data": "synthetic",
            "config": {"workload": desc, "channels_per_gpu": nch_default, "chunk_samples": chunk_len(w), "taps": ntaps,
                       "timed": "the reference's CPU implementation of the path (C port; Julia is not in the image) on ALL host "
                                "cores, rank 0 only at every N, each step a bounded channel SAMPLE of that workload (a rate, so "
                                "the sample size does not change the metric): " + r["sample"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def dtype_name(w):
    return {np.dtype(np.complex64): "c64 (f32 taps x complex64 samples, f32 FMA)", np.dtype(np.float32): "f32",
            np.dtype(np.float64): "f64", np.dtype(np.complex128): "c128"}[np.dtype(WORKLOADS[w][6])]


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Pin this rank's host threads (and therefore its first-touch pinned staging buffers) to the NUMA node its GPU
    hangs off: with one rank per GPU the end-to-end path otherwise funnels every rank's PCIe traffic through one
    socket's memory.  Best effort; returns the node or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class Ctx:
    """What every measurement needs: torch, the package, rank geometry, the clock sampler."""

    def __init__(self, args):
        import torch
        import multirate_b200 as mr
        self.torch, self.mr, self.args = torch, mr, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        self.numa = bind_to_gpu_numa(self.local) if self.world > 1 else None
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.sampler = ClockSampler(self.local)
        if self.rank == 0:
            self.sampler.start()                      # before any warm-up: short runs still get samples

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def reduce(self, vals, op):
        t = self.torch.tensor(vals, device="cuda", dtype=self.torch.float64)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def measure(ctx, w, steps, warmup, nch=None, e2e_steps=3, min_clock_window_s=0.4, policy=0):
    """Device-timed throughput of workload w (K steps bracketed by barrier + synchronize, CUDA events, max over ranks),
    roofline of its dominant kernel, and the end-to-end value through the public API with pinned host buffers."""
    torch, mr = ctx.torch, ctx.mr
    desc, ratio, ntaps, cutoff, beta, gain, tx, nch_default, nphi, po = WORKLOADS[w]
    nch = nch or nch_default
    n = chunk_len(w)
    f, h = make_filter(mr, w, nch, ctx.local)
    if policy:
        f.set_kernel_policy(policy)
    gen = torch.Generator(device="cuda"); gen.manual_seed(0x4D520000 + ctx.rank)
    if np.dtype(tx).kind == "c":
        x = torch.view_as_complex(torch.rand((nch, n, 2), generator=gen, device="cuda", dtype=torch.float32))
    else:
        x = torch.rand((nch, n), generator=gen, device="cuda",
                       dtype=torch.float64 if np.dtype(tx) == np.float64 else torch.float32)
    n_out_max = (f.outputlength(n) + 2 + 3) // 4 * 4                  # row pitch: a multiple of 16 bytes (TMA)
    ybuf = torch.empty((nch, n_out_max), dtype=x.dtype, device="cuda")
    es = x.element_size()
    small = nch * n * es <= (128 << 20)                               # working set could sit in the 126 MB L2
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def do_step():
        cnt = f._exact_count(n)
        f.filt_(ybuf, x)
        return cnt

    for _ in range(warmup):
        do_step()
    ctx.barrier()
    f.set_timing(True)
    l0 = f.launch_count
    outs = 0
    step_ms = []
    t_wall0 = time.time()
    if flush is None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(steps):
            outs += do_step() * nch
            ev[i + 1].record()
        ctx.barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    else:
        # small working set: L2 is flushed (256 MB written) before every step, outside the per-step events
        pairs = []
        torch.cuda.synchronize()
        for i in range(steps):
            flush.fill_(i & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            outs += do_step() * nch
            b.record()
            pairs.append((a, b))
        ctx.barrier()
        step_ms = [a.elapsed_time(b) for a, b in pairs]
        total_ms = float(sum(step_ms))
    t_wall1 = time.time()
    launches = f.launch_count - l0
    kms = f.kernel_ms()            # mean duration of the filt kernel(s) of one step, events on the launch stream
    kernel = f.last_kernel
    f.set_timing(False)
    clocks = None
    if ctx.rank == 0:
        label = "timed region"
        if t_wall1 - t_wall0 < min_clock_window_s:
            # too short for the 50 ms sampler: keep the same load running (untimed) and sample that
            t_end = time.time() + min_clock_window_s
            while time.time() < t_end:
                for _ in range(max(1, steps)):
                    do_step()
                torch.cuda.synchronize()
            t_wall1 = time.time()
            label = "timed region + untimed continuation of the same steps (region shorter than the sampling period)"
        clocks = ctx.sampler.window(t_wall0, t_wall1, label)
    elif t_wall1 - t_wall0 < min_clock_window_s:
        pass
    (total_ms,) = ctx.reduce([total_ms], "max")
    (outs_all,) = ctx.reduce([float(outs)], "sum")
    value = outs_all / (total_ms * 1e-3) / 1e6

    per_step_out = outs // steps
    alg_bytes = nch * n * es + per_step_out * es
    peak, peak_src = peaks()
    achieved = alg_bytes / (kms * 1e-3) / 1e9 if kms else None
    flops = 2 * TAPS_PER_OUT[w] * (2 if np.dtype(tx).kind == "c" else 1)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "kernel": kernel, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
    if kms:
        tf = per_step_out * flops / (kms * 1e-3) / 1e12
        fpeak = FP32_PEAK_TF / (2 if np.dtype(tx) == np.float64 else 1)
        roofline["fma"] = {"flops_per_output": flops, "achieved_tflops": tf, "peak_tflops_nominal": fpeak, "frac": tf / fpeak,
                           "unit": "TFLOP/s of useful FMA work (taps x outputs), against the nominal CUDA-core %s rate; a "
                                   "tensor-core kernel (mma_*) may exceed 1" % ("FP64" if np.dtype(tx) == np.float64 else "FP32")}
        b = BOUND[w]
        roofline["survey_bound"] = b
        if b in ("fp32", "fp64"):
            roofline["bound"] = "%s FMA per SURVEY 8d (see roofline.fma); hbm fields are the same launch against the HBM roof" % b
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % w)
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
        except Exception:
            pass

    e2e = None
    if e2e_steps > 0:
        xh_t = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        xh_t.copy_(x)
        yh_t = torch.empty((nch, n_out_max), dtype=x.dtype, pin_memory=True)
        xh, yh = xh_t.numpy(), yh_t.numpy()
        g, _ = make_filter(mr, w, nch, ctx.local)
        g.filt_(yh, xh)                                   # warm-up (allocates the staging buffers)
        ctx.barrier()
        t0 = time.perf_counter()
        eo = 0
        for _ in range(e2e_steps):
            r = g.filt_(yh, xh)
            eo += (r if isinstance(r, int) else per_step_out // nch) * nch
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        (dt,) = ctx.reduce([dt], "max")
        (eo,) = ctx.reduce([float(eo)], "sum")
        e2e = {"value": eo / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": nch * n * es,
               "d2h_bytes_per_step": int(per_step_out * es), "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "api": "FIRFilter.filt_(numpy pinned) -> mrb_filt_host", "numa_node": ctx.numa}
        hc = os.path.join(ROOT, "profiles", "host_ceiling.json")
        if os.path.exists(hc):
            try:
                c = json.load(open(hc))
                gbs = c.get("duplex_gbs", {}).get(str(ctx.world))
                if gbs:
                    e2e["host_ceiling_gbs"] = gbs
                    e2e["achieved_gbs"] = ctx.world and (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) * ctx.world / (dt / e2e_steps) / 1e9
                    e2e["host_ceiling_note"] = c.get("note")
            except Exception:
                pass
        del xh_t, yh_t, g
    res = {"value": value, "ms_per_step": total_ms / steps, "steps": steps, "step_ms_min_max": [min(step_ms), max(step_ms)],
           "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "channels_per_gpu": nch,
           "outputs_per_channel_per_step": per_step_out // nch,
           "l2": ("inputs (%.2f GiB per step per GPU) exceed the 126 MB L2; no flush needed" % (nch * n * es / 2 ** 30)) if not small
                 else "working set %.1f MiB per step: L2 flushed (256 MB written) before every timed step" % (nch * n * es / 2 ** 20)}
    del f, x, ybuf, flush
    torch.cuda.empty_cache()
    return res


def measure_c1_oneshot(ctx, reps=7):
    """The README benchmark as a user runs it (README.md:172-193): one-shot filt(h, x, 147//160) on HOST data --
    handle creation, H2D of 1e6 Float32 samples, kernel, D2H of 918,750 outputs, handle destruction -- wall clock."""
    mr, torch = ctx.mr, ctx.torch
    h = design_taps(3528, 0.5 / 147, 7.8562, 1.0)
    x = np.random.default_rng(1).random(1_000_000, dtype=np.float32)
    cold = []
    for _ in range(reps):                                      # every call builds and destroys its handle
        mr.clear_oneshot_cache()
        t0 = time.perf_counter()
        y = mr.filt(h, x, Fraction(147, 160))
        cold.append(time.perf_counter() - t0)
    mr.clear_oneshot_cache()
    ts = []
    for _ in range(reps):                                      # as a user repeats it: the handle of the last call is reset and reused
        t0 = time.perf_counter()
        y = mr.filt(h, x, Fraction(147, 160))
        ts.append(time.perf_counter() - t0)
    assert y.shape[0] == 918750
    med = float(np.median(ts[1:]))
    return {"seconds_median": med, "seconds_min": float(min(ts[1:])), "first_call_seconds": ts[0], "outputs": 918750,
            "seconds_median_new_handle_every_call": float(np.median(cold[1:])),
            "value": 918750 / med / 1e6, "unit": "Msamples/s",
            "readme_seconds": README_ONESHOT_S, "speedup_vs_readme": README_ONESHOT_S / med,
            "note": "README.md:190-193 ran Float64 taps on unspecified 2014 hardware, 1 thread; reported for scale only",
            "api": "filt(h, x, 147//160) on numpy Float32: mrb_create on the first call, then mrb_reset + mrb_filt_host on the handle kept for these taps "
                   "(seconds_median_new_handle_every_call: mrb_create + mrb_filt_host + mrb_destroy per call)",
            "h2d_bytes": 4_000_000, "d2h_bytes": 918750 * 4}


def measure_stream(ctx, log2n=31):
    """BASELINE configs[4], second half: ONE 2^31-sample complex64 stream through 147//160, split over the ranks into
    M-aligned input segments (segment_bounds): a segment that starts at a multiple of M starts from the constructor
    state, so rank r needs only the H samples before its segment (the halo) -- no collective, no state exchange.  Each
    rank filters its segment at multichannel speed (filt_long_stream: the segment viewed in place as a matrix of
    sub-segments).  Device-timed per rank, max over ranks."""
    torch, mr = ctx.torch, ctx.mr
    sharding = mr                        # segment_bounds / segment_plan / filt_long_stream are package exports
    L, M = 147, 160
    h = design_taps(3528, 0.5 / 147, 7.8562, 1.0)
    n_total = 1 << log2n
    n0, n1 = sharding.segment_bounds(n_total, ctx.world, ctx.rank, align=M * 2)
    H = 3528 // L - 1
    # synthetic stream: every segment is drawn from its own generator, its last H samples from a second one, so that
    # the next rank can draw its halo without receiving anything
    def gen_body(r, count):
        g = torch.Generator(device="cuda"); g.manual_seed(0x57000000 + r)
        return torch.view_as_complex(torch.rand((count, 2), generator=g, device="cuda", dtype=torch.float32))
    def gen_tail(r):
        g = torch.Generator(device="cuda"); g.manual_seed(0x58000000 + r)
        return torch.view_as_complex(torch.rand((H, 2), generator=g, device="cuda", dtype=torch.float32))
    nseg = n1 - n0
    x = torch.empty(nseg, dtype=torch.complex64, device="cuda")
    x[:nseg - H] = gen_body(ctx.rank, nseg - H)
    x[nseg - H:] = gen_tail(ctx.rank)
    halo0 = gen_tail(ctx.rank - 1) if ctx.rank > 0 else None
    plan = sharding.segment_plan(mr.FIRFilter(h, Fraction(L, M), device=-1), n_total, ctx.world, align=M * 2)
    ls = sharding.LongStream(h, Fraction(L, M), nseg, np.complex64, device=ctx.local)   # plan + handles: once per stream shape
    y = torch.empty(ls.total, dtype=torch.complex64, device="cuda")
    ls.run(x, halo0, out=y)                                                       # warm-up
    assert y.shape[0] == plan[ctx.rank][3], (y.shape, plan[ctx.rank])
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    a.record()
    ls.run(x, halo0, out=y)
    b.record()
    ctx.barrier()
    t1 = time.time()
    ms = a.elapsed_time(b)
    (ms_max,) = ctx.reduce([ms], "max")
    (outs,) = ctx.reduce([float(y.shape[0])], "sum")
    res = {"samples": n_total, "outputs": int(outs), "ms": ms_max, "value": outs / (ms_max * 1e-3) / 1e6, "unit": "Msamples/s",
           "segments": ctx.world, "halo_samples": H, "collective": "none",
           "algorithmic_gbs_per_gpu": (nseg + y.shape[0]) * 8 / (ms * 1e-3) / 1e9,
           "plan_rank0": {"n0": plan[0][0], "n1": plan[0][1], "k0": plan[0][2], "count": plan[0][3]},
           "note": "one LongStream.run per rank (halo gather, seek, launches; plan and handles built once); segment start states from the closed form "
                   "(k0 = ceil(n0 L / M), phase 0, deficit 1)"}
    if ctx.rank == 0:
        res["clocks"] = ctx.sampler.window(t0, t1)
    del x, y
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=0, help="override channels per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only-main", action="store_true", help="skip the other BASELINE configs (configs / stream / c1 one-shot)")
    ap.add_argument("--policy", type=int, default=0, help="1 = force the generic kernel, 2 = no tensor-core kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    ctx = Ctx(args)
    w = args.workload
    desc, ratio, ntaps, cutoff, beta, gain, tx, nch, nphi, po = WORKLOADS[w]
    m = measure(ctx, w, args.steps, args.warmup, nch=args.channels or None, e2e_steps=0 if args.no_e2e else max(1, min(args.steps, 3)),
                policy=args.policy)

    configs = {}
    oneshot = None
    stream = None
    if not args.only_main and w == "c5":
        for c in EXTRA_CONFIGS:
            # the other BASELINE configs: enough steps for a stable number (>= 20), weak-scaled like the headline
            try:
                r = measure(ctx, c, steps=20, warmup=3, e2e_steps=0 if args.no_e2e else 2, policy=args.policy)
                fr = r["roofline"]
                b = BOUND[c]
                frac = fr["frac"] if b == "hbm" else (fr.get("fma") or {}).get("frac") if b in ("fp32", "fp64") else None
                configs[c] = {"workload": WORKLOADS[c][0], "value": r["value"], "unit": "Msamples/s", "ms_per_step": r["ms_per_step"],
                              "kernel": fr["kernel"], "kernel_ms": fr["kernel_ms"], "bound": b, "frac": frac,
                              "hbm_frac": fr["frac"], "fma_frac_nominal": (fr.get("fma") or {}).get("frac"),
                              "e2e": r["e2e"], "clocks": r["clocks"], "gpu_launches": r["gpu_launches"], "l2": r["l2"]}
            except Exception as e:   # one config failing must not lose the headline
                configs[c] = {"error": repr(e)}
        try:
            stream = measure_stream(ctx)
        except Exception as e:
            stream = {"error": repr(e)}
        if ctx.rank == 0:
            try:
                oneshot = measure_c1_oneshot(ctx)
            except Exception as e:
                oneshot = {"error": repr(e)}
        ctx.barrier()

    if ctx.rank == 0:
        cpu = cpu1 = None
        if not args.no_cpu and ctx.world == 1:                 # the CPU baselines are timed at N = 1 only
            try:
                cpu = cpu_port_run(w, 3, 1)
                cpu.pop("ms_per_step", None)
                cpu1 = cpu_port_run(w, 2, 1, target_s=4.0, threads=1)   # the reference itself is single-threaded
                cpu1.pop("ms_per_step", None)
            except Exception as e:  # the baseline is a reported extra, never the product path
                cpu = cpu or {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        line = {"metric": "output Msamples/s (multichannel, device-timed)", "value": m["value"], "unit": "Msamples/s",
                "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(w),
                "data": "synthetic",
                "config": {"workload": desc, "channels_per_gpu": m["channels_per_gpu"], "chunk_samples": chunk_len(w), "taps": ntaps,
                           "outputs_per_channel_per_step": m["outputs_per_channel_per_step"], "l2": m["l2"],
                           "parallelism": "channels sharded over %d GPU(s), no collective" % ctx.world,
                           "step_ms_min_max": m["step_ms_min_max"]},
                "roofline": m["roofline"], "cpu_baseline": cpu, "cpu_baseline_1t": cpu1, "e2e": m["e2e"],
                "gpu_launches": m["gpu_launches"], "clocks": m["clocks"]}
        if configs:
            line["configs"] = configs
        if stream is not None:
            line["stream_2e31"] = stream
        if oneshot is not None:
            line["c1_oneshot"] = oneshot
        print(json.dumps(line), flush=True)
    ctx.sampler.stop()
    if ctx.dist:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
