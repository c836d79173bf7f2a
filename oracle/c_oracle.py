"""ctypes binding + on-demand build of oracle/mr_oracle.c.

TEST INFRASTRUCTURE ONLY (see mr_oracle.c).  Imported by tests/ and by
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mr_oracle.c")
BUILD = os.path.join(HERE, "_build")
KINDS = dict(standard=0, interpolator=1, decimator=2, rational=3, arbitrary=4, farrow=5)
DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}


def build(native: bool = False, outdir: str | None = None) -> str:
    """Compile mr_oracle.c.  `native=False` -> -march=x86-64-v3 (portable between the
    build container and the GPU box); `native=True` -> -march=native, for the timed
    baseline, built on the box that runs it."""
    outdir = outdir or BUILD
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, "libmr_oracle_native.so" if native else "libmr_oracle.so")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
        return out
    march = "-march=native" if native else "-march=x86-64-v3"
    tmp = out + ".tmp%d" % os.getpid()
    subprocess.check_call(["gcc", "-O3", march, "-fopenmp", "-fPIC", "-shared", "-std=gnu11",
                           "-o", tmp, SRC, "-lm"])
    os.replace(tmp, out)
    return out


_libs: dict = {}


def load(native: bool = False):
    if native in _libs:
        return _libs[native]
    try:
        path = build(native)
    except Exception:
        path = build(native, tempfile.mkdtemp(prefix="mro_"))
    lib = C.CDLL(path)
    lib.mro_create.restype = C.c_void_p
    lib.mro_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_double,
                               C.c_long, C.c_int, C.c_void_p, C.c_long]
    lib.mro_destroy.argtypes = [C.c_void_p]
    lib.mro_reset.argtypes = [C.c_void_p]
    lib.mro_outputlength.restype = C.c_long
    lib.mro_outputlength.argtypes = [C.c_void_p, C.c_long]
    lib.mro_filt.restype = C.c_long
    lib.mro_filt.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long, C.c_int]
    lib.mro_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.mro_max_threads.restype = C.c_int
    _libs[native] = lib
    return lib


class COracleFilter:
    """Multi-channel wrapper: x is (nch, n) C-contiguous; returns (nch, count)."""

    def __init__(self, kind, h, tx, nch, L=1, M=1, rate=0.0, Nphi=32, polyorder=0, pnfb=None, native=False):
        self.lib = load(native)
        h = np.asarray(h)
        self.th, self.tx = np.dtype(h.dtype), np.dtype(tx)
        self.ty = np.result_type(self.th, self.tx)
        hd = np.ascontiguousarray(h, dtype=np.float64)
        pn = None if pnfb is None else np.ascontiguousarray(pnfb, dtype=np.float64)
        self.nch = nch
        self.h = self.lib.mro_create(KINDS[kind], DT[self.th], DT[self.tx], hd.ctypes.data, len(hd), L, M, float(rate),
                                     Nphi, polyorder, None if pn is None else pn.ctypes.data, nch)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.mro_destroy(self.h)
            self.h = None

    def filt(self, x, nthreads=1, out=None):
        x = np.ascontiguousarray(x, dtype=self.tx)
        assert x.ndim == 2 and x.shape[0] == self.nch
        cap = max(int(self.lib.mro_outputlength(self.h, x.shape[1])), 0) + 2
        y = out if out is not None else np.empty((self.nch, cap), dtype=self.ty)
        n = self.lib.mro_filt(self.h, x.ctypes.data, x.shape[1], x.shape[1], y.ctypes.data, y.shape[1], nthreads)
        return y[:, :n]

    def reset(self):
        self.lib.mro_reset(self.h)

    def state(self):
        p, d = C.c_long(), C.c_long()
        a, al = C.c_double(), C.c_double()
        self.lib.mro_get_state(self.h, C.byref(p), C.byref(d), C.byref(a), C.byref(al))
        return dict(phiIdx=p.value, inputDeficit=d.value, acc=a.value, alpha=al.value)
