"""Multi-GPU partitioning of the path: channels are independent, so the batch is split into
contiguous channel blocks, one per GPU, with NO collective on the data path; a single long stream is split into
input segments that each need only a read-only tap-length halo and a closed-form start state (mrb_seek).
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
    count, from the library's own sequencing (host-only; `filt` must be an integer-ratio FIRFilter)."""
    from .filters import FIRFilter
    plan = []
    for r in range(world):
        n0, n1 = segment_bounds(n_samples, world, r, align)
        probe = FIRFilter(filt._h, filt._ratio, nchannels=1, sample_dtype="float32", device=-1)
        k0, cnt = C.c_int64(), C.c_int64()
        _ffi.check(_ffi.lib().mrb_seek(probe._handle, n0, None, 0, C.byref(k0), None))
        _ffi.check(_ffi.lib().mrb_output_count(probe._handle, n1 - n0, C.byref(cnt)))
        plan.append((n0, n1, k0.value, cnt.value))
    return plan
