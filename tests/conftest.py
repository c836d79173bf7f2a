import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _install_shared_farrow_fit():
    """The Farrow coefficients are an input of the filtering path and their least-squares fit is solver dependent at
    ~1e-10 (cond ~2.4e6): filtering parity uses ONE set of coefficients on both sides -- the library's agreed recipe,
    mrb_pfb2pnfb (host only, no GPU needed).  tests/test_farrow_fit.py removes the hook and pins that recipe against
    the oracle's own independent solve."""
    import multirate_b200 as mr
    import multirate_oracle as mo
    mo.set_pnfb_provider(lambda pfb, order: mr.pfb2pnfb(pfb, order))


_install_shared_farrow_fit()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def rand_samples(rng, shape, dtype):
    """rand(Tx, n) as the reference tests draw it (test/runtests.jl:398): U[0,1) (+ i U[0,1))."""
    dtype = np.dtype(dtype)
    x = rng.random(shape)
    if dtype.kind == "c":
        x = x + 1j * rng.random(shape)
    return x.astype(dtype)


def tol_for(dtype):
    """north_star tolerance: <= 1e-5 (Float32 / Complex64), <= 1e-12 (Float64 / Complex128), error
    normalised by max|y| ("Tolerance definition")."""
    return 1e-5 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else 1e-12


def nerr(y, ref):
    y, ref = np.asarray(y), np.asarray(ref)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return float(np.abs(y.astype(np.complex128) - ref.astype(np.complex128)).max() / max(np.abs(ref).max(), 1e-300))


@pytest.fixture
def rng():
    return np.random.default_rng(0x4D52)
