"""The Farrow polynomial fit (pfb2pnfb + polyfit, src/Filters.jl:311-321, src/support.jl:85-88) -- SURVEY 8 row a9.

The reference pins nothing here (test/farrowtest.jl only prints; Polynomials / `\\` are unpinned dependencies).  The
library's agreed recipe is mrb_pfb2pnfb: Householder QR of the Vandermonde matrix in Float64, coefficients rounded to the
tap type.  These tests pin that recipe
  * exactly, on data that IS a polynomial (the fit must return its coefficients),
  * against an independent solver (the oracle's numpy SVD least squares) with a stated bound -- the problem has
    cond ~2.4e6 (order 4) / ~1e8 (order 5), so two correct solvers differ around the 10th digit,
and bound what that solver dependence does to Farrow OUTPUTS (the reason filtering parity takes the coefficients as data).
Host only: no GPU needed."""
import numpy as np
import pytest

import multirate_b200 as mr
import multirate_oracle as mo


@pytest.fixture
def own_fit():
    """the oracle's independent fit for the duration of a test (conftest installs the shared one)"""
    mo.set_pnfb_provider(None)
    yield
    mo.set_pnfb_provider(lambda pfb, order: mr.pfb2pnfb(pfb, order))


def c4_taps(th):
    N = 32
    hLen, beta = mo.kaiserlength(0.05, samplerate=N)
    hLen = -(-hLen // N) * N
    return (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(th)


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5])
def test_fit_recovers_an_exact_polynomial(order):
    """Rows that are polynomials in phi = 1..Nphi with small integer coefficients: the least-squares solution is the
    polynomial itself, to rounding."""
    Nphi, T = 32, 7
    rng = np.random.default_rng(order)
    coef = rng.integers(-3, 4, size=(T, order + 1)).astype(np.float64)
    phi = np.arange(1, Nphi + 1, dtype=np.float64)
    pfb = np.stack([sum(coef[i, p] * phi ** p for p in range(order + 1)) for i in range(T)])       # (T, Nphi)
    got = mr.pfb2pnfb(pfb, order)
    assert got.shape == (T, order + 1)
    scale = np.abs(pfb).max(axis=1, keepdims=True)
    # residual of the fitted polynomial on the data, relative to the data
    fit = np.stack([sum(got[i, p] * phi ** p for p in range(order + 1)) for i in range(T)])
    assert np.abs(fit - pfb).max() <= 1e-9 * scale.max()
    assert np.abs(got - coef).max() <= 1e-6 * max(1.0, np.abs(coef).max())


@pytest.mark.parametrize("th", [np.float32, np.float64])
@pytest.mark.parametrize("order", [3, 4, 5])
def test_qr_fit_agrees_with_the_independent_svd_fit(th, order, own_fit):
    """mrb_pfb2pnfb (QR) against the oracle's numpy least squares (SVD) on the BASELINE configs[3] bank.  Bound: the
    coefficients of a row agree to 1e-8 of the row's largest coefficient (measured: 7e-11 at order 4, 4e-10 at order
    5); after rounding to Float32 taps at least 99 % of them are the same Float32 number."""
    h = c4_taps(th)
    a = mr.pfb2pnfb(mr.taps2pfb(h, 32), order)
    b = mo.pfb2pnfb(mo.taps2pfb(h, 32), order)
    assert a.shape == b.shape == (73, order + 1)
    scale = np.abs(b).max(axis=1, keepdims=True)
    assert (np.abs(a - b) / scale).max() <= 1e-8
    if th == np.float32:
        assert (a == b).mean() >= 0.99
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64))         # stored as Poly{Float32}


@pytest.mark.parametrize("th,bound", [(np.float32, 2e-7), (np.float64, 2e-9)])
def test_solver_dependence_of_farrow_outputs_is_bounded(th, bound, own_fit):
    """What the choice of solver does to the OUTPUTS: the oracle's Farrow filter run once with its own (SVD)
    coefficients and once with the library's (QR).  Float32 taps: the rounded coefficients are almost all identical, the
    outputs agree to ~1e-8; Float64: ~1e-10 -- above the 1e-12 filtering tolerance, which is why parity tests hand both
    sides the same coefficients."""
    h = c4_taps(th)
    x = np.random.default_rng(3).random(4000).astype(th)
    own = mo.FIRFilter(h, 0.918734, 32, 4)
    lib = mo.FIRFilter(h, 0.918734, 32, 4, pnfb=mr.pfb2pnfb(mr.taps2pfb(h, 32), 4))
    ya, yb = own.filt(x), lib.filt(x)
    assert ya.shape == yb.shape
    err = np.abs(ya.astype(np.float64) - yb.astype(np.float64)).max() / np.abs(ya).max()
    assert err <= bound, err
    assert own.state() == lib.state()                                            # sequencing does not depend on the taps


def test_create_without_coefficients_uses_the_library_fit():
    """mrb_create with poly_coeffs = NULL fits with the same recipe: tapsforphase agrees with the explicit route."""
    h = c4_taps(np.float32)
    f = mr.FIRFilter(h, 0.918734, 32, 4, device=-1, nchannels=1, sample_dtype=np.float32)
    g = mr.FIRFilter(h, 0.918734, 32, 4, device=-1, nchannels=1, sample_dtype=np.float32,
                     pnfb=mr.pfb2pnfb(mr.taps2pfb(h, 32), 4))
    for ph in (1.0, 7.25, 32.9):
        assert np.array_equal(mr.tapsforphase(f, ph), mr.tapsforphase(g, ph))
