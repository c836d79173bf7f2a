"""The C-ABI shared library loads here (no GPU) and exports every symbol include/mrb.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

import multirate_b200 as mr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mrb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    path = mr.build()
    assert os.path.exists(path)
    lib = mr._ffi.lib()
    assert b"sm_100a" in lib.mrb_version()


def test_every_declared_symbol_is_exported():
    syms = header_symbols()
    assert sorted(mr._ffi.SYMBOLS) == syms
    lib = ctypes.CDLL(mr._ffi.LIB)
    for s in syms:
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", mr._ffi.LIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mrb_\w+)", out))
    assert set(syms) <= exported


def test_library_holds_sm100a_code_only():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-lelf", mr._ffi.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_device_fails_loudly():
    """Without a usable GPU the compute entry points fail with a status; nothing falls back to the CPU."""
    import numpy as np
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    f = mr.FIRFilter(np.ones(8), nchannels=1, sample_dtype=np.float32, device=-1)
    with pytest.raises(mr.MrbError) as e:
        f.filt(np.ones(4, dtype=np.float32))
    assert e.value.code == mr._ffi.MRB_ERR_NO_DEVICE
    if not has_gpu:
        with pytest.raises(mr.MrbError) as e:
            mr.filt(np.ones(8), np.ones(4, dtype=np.float32))
        assert e.value.code in (mr._ffi.MRB_ERR_NO_DEVICE, mr._ffi.MRB_ERR_CUDA)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multirate.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "multirate_oracle" not in txt and "mr_oracle" not in txt and "c_oracle" not in txt, fn


def test_tiled_kernel_uses_uniform_datapath_taps_tma_and_ffma2():
    """The headline kernel's design rests on three SASS facts: taps arrive through the uniform
    datapath (LDCU, not vector LDC -- ptxas picks heuristically, so it is pinned here), samples and results move
    by TMA (UTMALDG / UTMASTG), and the complex x real FMA is a packed FFMA2."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-sass", mr._ffi.LIB], capture_output=True, text=True).stdout
    body = out[out.index("k_tiled_c64"):]
    nxt = body.find("Function :", 10)
    body = body[:nxt] if nxt > 0 else body
    n_ldcu = len(re.findall(r"\bLDCU(\.64)? UR\d+, c\[0x0\]\[UR", body))
    n_ldc_vec = len(re.findall(r"\bLDC(\.64)? R\d+, c\[0x0\]\[R", body))
    assert len(re.findall(r"\bFFMA2\b", body)) >= 288
    assert n_ldcu >= 140 and n_ldc_vec <= 8, (n_ldcu, n_ldc_vec)
    assert "UTMALDG" in body and "UTMASTG" in body
