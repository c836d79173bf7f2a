#!/usr/bin/env python
"""bench.py -- streaming polyphase FIR throughput on B200 (the driver's contract, tier ④ reading).

A "step" is one pass of the hot path over one batch of synthetic input: one 65,536-sample chunk per channel
through a stateful FIRFilter (history, phase and deficit carried on the device between steps).

Default workload (N=1 and every N, weak scaling): BASELINE.json configs[4]'s per-GPU shard --
FIRRational 147//160, 3528-tap Kaiser low-pass (Float32 taps), 8192 channels of Complex64 per GPU
(65,536 channels over 8 GPUs), 64K-sample chunks.  Other configs: --workload c1|c2|c3a|c3b|c4a|c4f|c4a64|c4f64 (and x* extras).

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
    # extras (not BASELINE configs): the same kernels on the other sample type / the mirrored ratio
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
}


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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


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
# CPU baseline: the C restatement of the reference loops (oracle/mr_oracle.c), all host threads.
# ------------------------------------------------------------------------------------------------
def cpu_port_run(w, steps, warmup, target_s=6.0):
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
    threads = os.cpu_count() or 1
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
                      "(oracle/mr_oracle.c, gcc -O3 %s, OpenMP over channels); Julia reference not runnable: no julia in image"
                      % (want, n, steps, "-march=native" if native else "-march=x86-64-v3"),
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
                       "timed": "bounded sample of that workload on the host cores: " + r["sample"]},
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
    ap.add_argument("--policy", type=int, default=0, help="1 = force the generic kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import multirate_b200 as mr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    numa = bind_to_gpu_numa(local) if world > 1 else None
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = args.workload
    desc, ratio, ntaps, cutoff, beta, gain, tx, nch, nphi, po = WORKLOADS[w]
    nch = args.channels or nch
    n = chunk_len(w)
    f, h = make_filter(mr, w, nch, local)
    if args.policy:
        f.set_kernel_policy(args.policy)
    # synthetic samples U[0,1) (+ i U[0,1)) generated on the device, seed = 0x4D520000 + rank
    gen = torch.Generator(device="cuda"); gen.manual_seed(0x4D520000 + rank)
    if np.dtype(tx).kind == "c":
        x = torch.view_as_complex(torch.rand((nch, n, 2), generator=gen, device="cuda", dtype=torch.float32))
    else:
        x = torch.rand((nch, n), generator=gen, device="cuda",
                       dtype=torch.float64 if np.dtype(tx) == np.float64 else torch.float32)
    n_out_max = (f.outputlength(n) + 2 + 3) // 4 * 4                  # row pitch: a multiple of 16 bytes (TMA)
    ybuf = torch.empty((nch, n_out_max), dtype=x.dtype, device="cuda")
    es = x.element_size()

    # one step: filt! into a preallocated device buffer; returns the per-channel output count
    def do_step():
        cnt = f._exact_count(n)
        f.filt_(ybuf, x)
        return cnt

    for _ in range(args.warmup):
        do_step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    f.set_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = f.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    outs = 0
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(args.steps):
        outs += do_step() * nch
        ev[i + 1].record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches = f.launch_count - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, float(outs)], device="cuda", dtype=torch.float64)
    if dist:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, outs_all = tmax[0].item(), tsum[1].item()
    else:
        outs_all = float(outs)
    value = outs_all / (total_ms * 1e-3) / 1e6

    # roofline of the dominant kernel: algorithmic bytes per launch / its own CUDA-event duration
    kms = f.kernel_ms()            # mean duration of the filt kernel(s) of one step, events on the launch stream
    per_step_out = outs // args.steps
    alg_bytes = nch * n * es + per_step_out * es
    peak, peak_src = peaks()
    achieved = alg_bytes / (kms * 1e-3) / 1e9 if kms else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "kernel": f.last_kernel, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src}
    # FP32 side of the roofline: real FMAs the kernel executes per output and channel.  Only the 147//160
    # shard (c5) is HBM-bound; decimator-256, standard-128 and the arbitrary-rate kernels sit on the FP32 roof.
    taps_per_out = {"c5": 24, "c1": 24, "c2": 256, "c3a": 32, "c3b": 128, "c4a": 73, "c4f": 73, "c4a64": 73, "c4f64": 73,
                    "x160": 24, "x2f": 256, "x3ac": 32, "x3bc": 128, "x4ac": 73, "x4fc": 73}[w]
    flops = 2 * taps_per_out * (2 if np.dtype(tx).kind == "c" else 1)
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12                       # nominal, TFLOP/s
    if kms:
        tf = per_step_out * flops / (kms * 1e-3) / 1e12
        roofline["fp32"] = {"flops_per_output": flops, "achieved_tflops": tf, "peak_tflops_nominal": fp32_peak,
                            "frac": tf / fp32_peak,
                            "note": "arbitrary: taps blended once per output (73 FMAs), the reference does two dot products"
                                    if w == "c4a" else None}
        if w.startswith("c4") and w.endswith("64"):
            roofline["fp32"]["note"] = "Float64 FMAs; the nominal FP64 peak is half the FP32 figure"
        if w in ("c2", "c3b", "c4a", "c4f", "c4a64", "c4f64", "x2f", "x3bc", "x4ac", "x4fc"):
            roofline["bound"] = "fp32 (see roofline.fp32; hbm fields kept for reference)"
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % w)
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 3))
        xh_t = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        xh_t.copy_(x)
        yh_t = torch.empty((nch, n_out_max), dtype=x.dtype, pin_memory=True)
        xh, yh = xh_t.numpy(), yh_t.numpy()
        g, _ = make_filter(mr, w, nch, local)
        g.filt_(yh, xh)                                   # warm-up (allocates the staging buffers)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        eo = 0
        for _ in range(e2e_steps):
            r = g.filt_(yh, xh)
            eo += (r if isinstance(r, int) else per_step_out // nch) * nch
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, float(eo)], device="cuda", dtype=torch.float64)
        if dist:
            a = tt.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = tt.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            dt, eo = a[0].item(), b[1].item()
        e2e = {"value": eo / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": nch * n * es,
               "d2h_bytes_per_step": int(per_step_out * es), "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "api": "FIRFilter.filt_(numpy pinned) -> mrb_filt_host", "numa_node": numa}
        del xh_t, yh_t

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:                 # the CPU baseline is timed at N = 1 only
            try:
                cpu = cpu_port_run(w, 3, 1)
                cpu.pop("ms_per_step", None)
            except Exception as e:  # the baseline is a reported extra, never the product path
                cpu = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        line = {"metric": "output Msamples/s (multichannel, device-timed)", "value": value, "unit": "Msamples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(w),
                "data": "synthetic",
                "config": {"workload": desc, "channels_per_gpu": nch, "chunk_samples": n, "taps": ntaps,
                           "outputs_per_channel_per_step": per_step_out // nch,
                           "l2": "inputs (%.2f GiB per step per GPU) exceed the 126 MB L2; no flush needed"
                                 % (nch * n * es / 2 ** 30) if nch * n * es > (512 << 20) else
                                 "working set %.1f MiB per step" % (nch * n * es / 2 ** 20),
                           "parallelism": "channels sharded over %d GPU(s), no collective" % world,
                           "step_ms_min_max": [min(step_ms), max(step_ms)]},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
