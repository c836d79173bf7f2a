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


class LongStream:
    """One very long single-channel stream at multichannel speed, with no copy and no collective (BASELINE
    configs[4] "one 2^31-sample stream split into segments with tap-length halo").

    The stream (or one rank's M-aligned segment of it) is viewed in place as a channel-major matrix of `rows` sub-segments
    of `seg` samples, seg a multiple of the decimation M: every sub-segment then starts from the constructor state (phase 0,
    deficit 1; src/Filters.jl:567-571 in closed form), its history is simply the H samples that precede it in memory (the
    halo), and it produces exactly seg*L/M outputs, so the output matrix is the output stream, in place.  The tail that
    does not fill a sub-segment runs through a one-channel filter seeked to its position.

    The plan and the two handles are built once for a stream length `n` and sample dtype; `run(x, halo0)` then filters any
    stream of that length (e.g. successive blocks of a recording, or the same block on every step of a benchmark):
    `halo0` = the H samples that precede x[0] (None at the start of the stream).  Across GPUs every rank builds one for its
    own segment (`segment_bounds(..., align=M)`); nothing but the halo crosses between ranks."""

    def __init__(self, h, ratio, n, dtype, device=0, rows_target: int = 8192):
        import math
        from fractions import Fraction

        import numpy as np

        from .filters import FIRFilter
        self.h, self.ratio = h, Fraction(ratio)
        L, M = self.ratio.numerator, self.ratio.denominator
        self.n, self.tx = int(n), np.dtype(dtype)
        probe = FIRFilter(h, self.ratio, nchannels=1, sample_dtype=self.tx, device=-1)
        self.H = probe.historyLen
        # sub-segment length: a multiple of M whose input and output row pitches are multiples of 16 bytes
        al = max(1, 16 // self.tx.itemsize)
        unit = M * al // math.gcd(M, al)
        while (unit * L // M) % al:
            unit *= 2
        self.seg = max(unit, -(-(-(-self.n // max(rows_target, 1))) // unit) * unit)     # ceil: at most rows_target rows
        self.rows = self.n // self.seg
        self.seg_out = self.seg * L // M
        self.total = probe._exact_count(self.n)
        self.body = self.rows >= 2 and self.seg > 4 * self.H
        self.done_in = self.rows * self.seg if self.body else 0
        self.done_out = self.rows * self.seg_out if self.body else 0
        self.f = FIRFilter(h, self.ratio, nchannels=self.rows, sample_dtype=self.tx, device=device) if self.body else None
        self.g = FIRFilter(h, self.ratio, nchannels=1, sample_dtype=self.tx, device=device) if self.done_in < self.n else None
        self._halo = None
        self._idx = None

    def run(self, x, halo0=None, out=None):
        import torch
        assert x.dim() == 1 and x.shape[0] == self.n
        H, lib = self.H, _ffi.lib()
        y = out if out is not None else torch.empty(self.total, dtype=x.dtype, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        k0 = C.c_int64()
        if self.body:
            body = x[:self.rows * self.seg].view(self.rows, self.seg)
            if self._halo is None:
                self._halo = torch.zeros((self.rows, max(H, 1)), dtype=x.dtype, device=x.device)
                if H:
                    self._idx = (torch.arange(1, self.rows, device=x.device) * self.seg).unsqueeze(1) + torch.arange(-H, 0, device=x.device)
            if H:
                self._halo[1:, :H] = x[self._idx]
                if halo0 is not None:
                    self._halo[0, :H] = halo0[-H:]
                else:
                    self._halo[0].zero_()
            _ffi.check(lib.mrb_seek(self.f._handle, 0, self._halo.data_ptr() if H else None, max(H, 1), C.byref(k0), stream))
            self.f.filt_(y[:self.rows * self.seg_out].view(self.rows, self.seg_out), body)
        if self.g is not None:
            halo = x[self.done_in - H:self.done_in].contiguous() if (self.done_in and H) else (
                halo0[-H:].contiguous() if (halo0 is not None and H) else None)
            _ffi.check(lib.mrb_seek(self.g._handle, self.done_in, halo.data_ptr() if halo is not None else None, H, C.byref(k0), stream))
            assert k0.value == self.done_out
            self.g.filt_(y[self.done_out:], x[self.done_in:])
        return y


def filt_long_stream(h, ratio, x, rows_target: int = 8192, halo0=None):
    """One-call form of LongStream (plan and handles built per call).  x: 1-D CUDA tensor (torch); returns the 1-D output
    tensor.  Across GPUs the same call filters ONE RANK'S SEGMENT of the stream: give every rank a segment that starts at a
    multiple of M (`segment_bounds(..., align=M)`) and pass `halo0` = the H samples that precede it (None for the first)."""
    import numpy as np
    tx = np.dtype(str(x.dtype).replace("torch.", ""))
    return LongStream(h, ratio, x.shape[0], tx, device=x.device.index or 0, rows_target=rows_target).run(x, halo0)
