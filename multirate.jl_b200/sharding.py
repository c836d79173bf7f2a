"""Multi-GPU partitioning of the path: channels are independent, so the batch is split into
contiguous channel blocks, one per GPU, with NO collective on the data path; a single long stream is split into
input segments that each need only a read-only tap-length halo and a start state (mrb_seek: closed form for integer
ratios, exact phase replay for arbitrary / Farrow).
Host-side arithmetic only."""
from __future__ import annotations

import ctypes as C

from . import _ffi


def channel_shard(n_channels: int, world: int, rank: int):
    """Contiguous block of channels for `rank`: [lo, hi). Blocks differ in size by at most one channel."""
    base, rem = divmod(n_channels, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def segment_bounds(n_samples: int, world: int, rank: int, align: int = 1):
    """Input segment [lo, hi) of a long stream for `rank`; interior boundaries are multiples of `align`
    (choosing align = decimation makes every interior segment start at phase 0 with deficit 1)."""
    per = -(-n_samples // world)
    per = -(-per // align) * align
    lo = min(rank * per, n_samples)
    return lo, min(lo + per, n_samples)


def segment_plan(filt, n_samples: int, world: int, align: int = 1):
    """[(n0, n1, k0, count)] for every rank: segment bounds, absolute index of its first output, and its output
    count, from the library's own sequencing (host-only).  Integer ratios use the closed-form start state; arbitrary
    and Farrow filters the exact replay of their phase recurrence, O(n0) host work per segment."""
    from .filters import FIRFilter
    plan = []
    for r in range(world):
        n0, n1 = segment_bounds(n_samples, world, r, align)
        probe = FIRFilter(*filt._ctor_args(), nchannels=1, sample_dtype="float32", device=-1)
        k0, cnt = C.c_int64(), C.c_int64()
        _ffi.check(_ffi.lib().mrb_seek(probe._handle, n0, None, 0, C.byref(k0), None))
        _ffi.check(_ffi.lib().mrb_output_count(probe._handle, n1 - n0, C.byref(cnt)))
        plan.append((n0, n1, k0.value, cnt.value))
    return plan


def filt_long_stream(h, ratio, x, rows_target: int = 8192, halo0=None):
    """One very long single-channel stream at multichannel speed, with no copy and no collective (SURVEY 8e,
    BASELINE configs[4] "one 2^31-sample stream split into segments with tap-length halo").

    The stream is viewed in place as a channel-major matrix of `rows` segments of `seg` samples, seg a multiple of the
    decimation M: every segment then starts from the constructor state (phase 0, deficit 1; src/Filters.jl:567-571
    in closed form), its history is simply the H samples that precede it in memory (the halo), and it produces exactly
    seg*L/M outputs, so the output matrix is the output stream, in place.  The tail that does not fill a segment runs
    through a one-channel filter seeked to its position.  x: 1-D CUDA tensor (torch); returns the 1-D output tensor.

    Across GPUs the same call filters ONE RANK'S SEGMENT of the stream: give every rank a segment that starts at a
    multiple of M (`segment_bounds(..., align=M)`) and pass `halo0` = the H samples that precede it (None for the
    first segment).  A segment that starts at a multiple of M starts from the constructor state, so its outputs are
    the stream's outputs [n0*L/M, ...) and no state or sample crosses between GPUs at run time.
    """
    import math
    from fractions import Fraction

    import numpy as np
    import torch

    from .filters import FIRFilter
    ratio = Fraction(ratio)
    L, M = ratio.numerator, ratio.denominator
    n = x.shape[0]
    tx = np.dtype(str(x.dtype).replace("torch.", ""))
    probe = FIRFilter(h, ratio, nchannels=1, sample_dtype=tx, device=-1)
    H = probe.historyLen
    # segment length: a multiple of M whose input and output row pitches are multiples of 16 bytes
    al = max(1, 16 // tx.itemsize)
    unit = M * al // math.gcd(M, al)
    while (unit * L // M) % al:
        unit *= 2
    seg = max(unit, -(-(-(-n // max(rows_target, 1))) // unit) * unit)       # ceil: at most rows_target segments
    rows = n // seg
    seg_out = seg * L // M
    total = probe._exact_count(n)
    y = torch.empty(total, dtype=x.dtype, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    lib = _ffi.lib()
    done_in = done_out = 0
    if rows >= 2 and seg > 4 * H:
        body = x[:rows * seg].view(rows, seg)
        f = FIRFilter(h, ratio, nchannels=rows, sample_dtype=tx, device=x.device.index or 0)
        halo = torch.zeros((rows, max(H, 1)), dtype=x.dtype, device=x.device)
        if H:
            idx = (torch.arange(1, rows, device=x.device) * seg).unsqueeze(1) + torch.arange(-H, 0, device=x.device)
            halo[1:, :H] = x[idx]
            if halo0 is not None:
                halo[0, :H] = halo0[-H:]
        k0 = C.c_int64()
        _ffi.check(lib.mrb_seek(f._handle, 0, halo.data_ptr() if H else None, max(H, 1), C.byref(k0), stream))
        f.filt_(y[:rows * seg_out].view(rows, seg_out), body)
        done_in, done_out = rows * seg, rows * seg_out
    if done_in < n:
        g = FIRFilter(h, ratio, nchannels=1, sample_dtype=tx, device=x.device.index or 0)
        k0 = C.c_int64()
        halo = x[done_in - H:done_in].contiguous() if (done_in and H) else (
            halo0[-H:].contiguous() if (halo0 is not None and H) else None)
        _ffi.check(lib.mrb_seek(g._handle, done_in, halo.data_ptr() if halo is not None else None, H, C.byref(k0), stream))
        assert k0.value == done_out
        g.filt_(y[done_out:], x[done_in:])
    return y
