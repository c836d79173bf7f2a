"""ctypes binding of libmrb (include/mrb.h) + the in-tree nvcc build recipe.

The library is the product's only compute path: if it is missing or cannot be
loaded this module raises -- there is no CPU fallback and nothing here imports
the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmrb.so")
SOURCES = ["mrb_api.cu"]
HEADERS = ["mrb_kernels.cuh", "mrb_tiled.cuh", "mrb_unit.cuh", "mrb_decim.cuh", "mrb_table.cuh", "mrb_mma.cuh", "mrb_seq.h", os.path.join("..", "..", "include", "mrb.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC,-ffp-contract=off"]

MRB_OK, MRB_ERR_BAD_ARGUMENT, MRB_ERR_BUFFER_TOO_SMALL, MRB_ERR_CUDA, MRB_ERR_UNSUPPORTED, MRB_ERR_NO_DEVICE = range(6)
KIND_AUTO, STANDARD, INTERPOLATOR, DECIMATOR, RATIONAL, ARBITRARY, FARROW = -1, 0, 1, 2, 3, 4, 5
F32, F64, C64, C128 = 0, 1, 2, 3

# every symbol include/mrb.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = ["mrb_create", "mrb_destroy", "mrb_get_info", "mrb_outputlength", "mrb_output_count", "mrb_inputlength",
           "mrb_nextphase", "mrb_taps2pfb", "mrb_pfb2pnfb", "mrb_filt", "mrb_filt_host", "mrb_set_host_pipeline", "mrb_advance", "mrb_reset", "mrb_setphase",
           "mrb_get_state", "mrb_set_state", "mrb_get_history", "mrb_set_history", "mrb_tapsforphase", "mrb_get_pfb",
           "mrb_seek", "mrb_get_schedule", "mrb_set_taps", "mrb_set_taps_async", "mrb_launch_count", "mrb_set_timing", "mrb_get_timing", "mrb_set_kernel_policy", "mrb_last_kernel", "mrb_last_error", "mrb_version"]


class Desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tap_dtype", C.c_int32), ("sample_dtype", C.c_int32), ("device", C.c_int32),
                ("h", C.c_void_p), ("h_len", C.c_int64), ("interpolation", C.c_int64), ("decimation", C.c_int64),
                ("rate", C.c_double), ("n_phi", C.c_int32), ("poly_order", C.c_int32), ("poly_coeffs", C.c_void_p),
                ("n_channels", C.c_int64)]


class State(C.Structure):
    _fields_ = [("phi_idx", C.c_int64), ("input_deficit", C.c_int64), ("x_idx", C.c_int64),
                ("phi_accumulator", C.c_double), ("alpha", C.c_double)]


class Info(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tap_dtype", C.c_int32), ("sample_dtype", C.c_int32), ("out_dtype", C.c_int32),
                ("device", C.c_int32), ("n_phi", C.c_int32), ("poly_order", C.c_int32),
                ("taps_per_phase", C.c_int64), ("history_len", C.c_int64), ("h_len", C.c_int64),
                ("interpolation", C.c_int64), ("decimation", C.c_int64), ("n_channels", C.c_int64), ("rate", C.c_double)]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libmrb.so (in-tree, travels with gpurun)."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    return LIB


_lib = None


def lib():
    """Load libmrb.so; raise loudly when it is absent (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            build()
        else:
            raise RuntimeError("libmrb.so is not built (%s) and nvcc is unavailable; the CUDA extension is "
                               "the only compute path of this package" % LIB)
    L = C.CDLL(LIB)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "mrb_create": (i32, [P(Desc), P(vp)]), "mrb_destroy": (i32, [vp]), "mrb_get_info": (i32, [vp, P(Info)]),
        "mrb_outputlength": (i32, [vp, i64, P(i64)]), "mrb_output_count": (i32, [vp, i64, P(i64)]),
        "mrb_inputlength": (i32, [i64, i64, i64, i64, P(i64)]), "mrb_nextphase": (i32, [i64, i64, i64, P(i64)]),
        "mrb_taps2pfb": (i32, [vp, i64, i32, i64, vp]),
        "mrb_pfb2pnfb": (i32, [vp, i64, i32, i64, i32, vp]),
        "mrb_filt": (i32, [vp, vp, i64, i64, vp, i64, i64, P(i64), vp]),
        "mrb_filt_host": (i32, [vp, vp, i64, i64, vp, i64, i64, P(i64)]),
        "mrb_set_host_pipeline": (i32, [vp, i32, i32]),
        "mrb_advance": (i32, [vp, i64, P(i64)]), "mrb_reset": (i32, [vp]), "mrb_setphase": (i32, [vp, dbl]),
        "mrb_get_state": (i32, [vp, P(State)]), "mrb_set_state": (i32, [vp, P(State)]),
        "mrb_get_history": (i32, [vp, vp]), "mrb_set_history": (i32, [vp, vp]),
        "mrb_tapsforphase": (i32, [vp, dbl, vp]), "mrb_get_pfb": (i32, [vp, i32, vp]),
        "mrb_seek": (i32, [vp, i64, vp, i64, P(i64), vp]), "mrb_set_taps": (i32, [vp, vp, i64, vp]), "mrb_set_taps_async": (i32, [vp, vp, i64, vp, vp]), "mrb_get_schedule": (i32, [vp, i64, vp, vp, vp]), "mrb_launch_count": (i32, [vp, P(i64)]),
        "mrb_set_kernel_policy": (i32, [vp, i32]),
        "mrb_set_timing": (i32, [vp, i32]), "mrb_get_timing": (i32, [vp, P(dbl), P(i64)]), "mrb_last_kernel": (C.c_char_p, [vp]),
        "mrb_last_error": (C.c_char_p, []), "mrb_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


class MrbError(RuntimeError):
    """Non-zero status from libmrb; `.code` is the mrb_status, the message is mrb_last_error()
    (the reference's wording where it has one, e.g. "buffer is too small", src/Filters.jl:550)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def check(rc: int):
    if rc != MRB_OK:
        raise MrbError(rc, lib().mrb_last_error().decode())
