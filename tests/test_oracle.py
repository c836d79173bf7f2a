"""Pins the oracle (oracle/multirate_oracle.py, oracle/mr_oracle.c) on every known-answer vector the
reference holds for the path and on the textbook definition its own tests use.  CPU only."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

import c_oracle as co
import multirate_oracle as mo
from conftest import nerr, rand_samples, tol_for

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "kat.json")))


def test_readme_3_17_kat():
    """README.md:58-142: values per chunk, printed pfb, initial struct, chunked == one-shot."""
    k = KAT["readme_3_17"]
    h, x = np.array(k["h"], dtype=np.float64), np.array(k["x"], dtype=np.float64)
    f = mo.FIRFilter(h, Fraction(*k["ratio"]))
    assert isinstance(f.kernel, mo.FIRRational)
    assert np.array_equal(f.kernel.pfb, np.array(k["pfb"]))
    assert (f.kernel.Nphi, f.kernel.tapsPerphi, f.kernel.criticalYidx, f.kernel.phiIdx, f.kernel.inputDeficit) == \
        (k["Nphi"], k["tapsPerphi"], k["criticalYidx"], k["phiIdx"], k["inputDeficit"])
    assert f.historyLen == k["historyLen"]
    pos, ys = 0, []
    for n, want in zip(k["chunks"], k["y"]):
        y = f.filt(x[pos:pos + n]); pos += n
        assert np.array_equal(y, np.array(want))
        ys.append(y)
    assert np.sum(np.concatenate(ys) - mo.filt(h, x, Fraction(*k["ratio"]))) == 0.0      # README.md:140-141


def test_taps2pfb_example():
    k = KAT["taps2pfb_example"]
    assert np.array_equal(mo.taps2pfb(np.array(k["h"]), k["Nphi"]), np.array(k["pfb"]))


def test_nextphase():
    """test/runtests.jl:423-438."""
    for interpolation in range(1, 9):
        for decimation in range(1, 9):
            ratio = Fraction(interpolation, decimation)
            L, M = ratio.numerator, ratio.denominator
            x = np.tile(np.arange(1, L + 1), M)
            reference = x[::M]
            result = [1]
            for _ in range(2, L + 1):
                result.append(mo.nextphase(result[-1], ratio))
            assert list(reference) == result


def test_farrow_notebook_count():
    k = KAT["farrow_notebook_count"]
    N = k["Nphi"]
    h = mo.firdes(k["tapsPerphi"] * N, min(0.45 / N, k["rate"] / N)) * N
    t = np.arange(k["n_in"])
    x = np.cos(2 * np.pi * 0.15 * t) + 0.5 * np.sin(2 * np.pi * 0.3 * t * np.pi)
    assert len(mo.filt(h, x, k["rate"], N, k["polyorder"])) == k["n_out"]


def test_readme_benchmark_count():
    k = KAT["readme_benchmark_count"]
    f = mo.FIRFilter(np.ones(3528), Fraction(*k["ratio"]))
    assert f.outputlength(k["n_in"]) == k["n_out"] and k["n_out"] * 8 <= k["bytes"]


@pytest.mark.parametrize("th", [np.float32, np.float64])
@pytest.mark.parametrize("tx", [np.float32, np.float64, np.complex64, np.complex128])
def test_four_way_equivalence(th, tx, rng):
    """test/runtests.jl:46-324: naive definition == one-shot == 2-chunk == sample-at-a-time, for single
    rate, decimation, interpolation and rational ratios (seeded restatement of test_all :389-421)."""
    Ls = [1] + sorted(set(rng.integers(2, 33, 3).tolist()))
    Ms = [1] + sorted(set(rng.integers(2, 33, 3).tolist()))
    for L in Ls:
        for M in Ms:
            ratio = Fraction(L, M)
            h = rng.random(int(rng.integers(16, 129))).astype(th)
            xLen = int(rng.integers(200, 301)); xLen -= xLen % M
            x = rand_samples(rng, xLen, tx)
            naive = mo.naivefilt(h, x, ratio)
            one = mo.filt(h, x, ratio)
            f = mo.FIRFilter(h, ratio)
            piv = min(int(rng.integers(50, 151)), xLen // 4)
            two = np.concatenate([f.filt(x[:piv]), f.filt(x[piv:])])
            f.reset()
            pw = np.concatenate([f.filt(x[i:i + 1]) for i in range(xLen)])
            tol = 2 * tol_for(np.result_type(th, tx))
            assert len(one) == len(naive)
            assert nerr(one, naive.astype(one.dtype)) < max(tol, 1e-6 if np.dtype(th) == np.float32 else 0)
            assert nerr(two, one) < tol and nerr(pw, one) < tol
            # c restatement (native-precision accumulation, like the reference)
            kind = "standard" if ratio == 1 else "decimator" if ratio.numerator == 1 else \
                "interpolator" if ratio.denominator == 1 else "rational"
            c = co.COracleFilter(kind, h, tx, 1, ratio.numerator, ratio.denominator)
            yc = np.concatenate([c.filt(x[None, :piv])[0], c.filt(x[None, piv:])[0]])
            assert nerr(yc, one) < 20 * tol_for(np.result_type(th, tx))
            st, so = c.state(), f.state()
            # after the piecewise run the python oracle consumed the same xLen samples
            assert all(st[k] == so[k] for k in so)


@pytest.mark.parametrize("tx", [np.float32, np.complex64])
def test_arbitrary_vs_naive_and_piecewise(tx, rng):
    """test/runtests.jl:335-378: loose comparison with NaiveResamplers (truncated to the common length)
    and exact chunking invariance of counts and state."""
    N = 32
    hLen, beta = mo.kaiserlength(0.05, samplerate=N)
    hLen = -(-hLen // N) * N
    h = (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(np.float32)
    assert hLen == 2336
    x = rand_samples(rng, 257, tx)
    for rate in (0.918734, 1.0 + 0.3712, 2.5):
        naive = mo.naivefilt_arbitrary(h.astype(np.float64), x, rate, N)
        one = mo.filt(h, x, rate, N)
        f = mo.FIRFilter(h, rate, N)
        pw = np.concatenate([f.filt(x[i:i + 1]) for i in range(len(x))])
        assert len(pw) == len(one) and nerr(pw, one) < 1e-6
        n = min(len(naive), len(one))
        assert abs(len(naive) - len(one)) <= 1
        assert np.abs(naive[:n] - one[:n]).max() < 5e-3 * max(1.0, np.abs(one).max())
        g = mo.FIRFilter(h, rate, N); g.filt(x)
        assert g.state() == f.state()


def test_farrow_piecewise_and_close_to_arbitrary(rng):
    N = 32
    h = (mo.firdes(320, 0.45 / N) * N)
    x = rng.random(300)
    one = mo.filt(h, x, 0.918734, N, 4)
    f = mo.FIRFilter(h, 0.918734, N, 4)
    pw = np.concatenate([f.filt(x[i:i + 1]) for i in range(len(x))])
    assert len(pw) == len(one) and nerr(pw, one) < 1e-13
    arb = mo.filt(h, x, 0.918734, N)
    assert len(arb) == len(one) and nerr(one, arb) < 5e-2          # test/farrowtest.jl only prints this


def test_c_oracle_matches_golden_vectors():
    """The frozen oracle vectors (tests/golden/oracle_vectors.npz) against the C restatement."""
    g = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
    names = sorted({k.rsplit(".", 1)[0] for k in g.files})
    assert len(names) == 56
    for key in names:
        name, th, tx = key.split(".")
        h, x = g[key + ".h"], g[key + ".x"]
        kind = name.split("_")[0]
        L, M = {"standard": (1, 1), "decimator": (1, 8), "interpolator": (4, 1), "rational": (147, 160),
                "rational_3_17": (3, 17)}.get(name, (1, 1))
        pn = mo.pfb2pnfb(mo.taps2pfb(h, 32), 4) if kind == "farrow" else None
        c = co.COracleFilter(kind, h, x.dtype, 2, L, M, rate=0.918734, Nphi=32, polyorder=4, pnfb=pn)
        for i, (a, b) in enumerate(((0, 1), (1, 40), (40, 331))):
            y = c.filt(x[:, a:b], nthreads=2)
            assert nerr(y, g[key + ".y%d" % i]) < 20 * tol_for(y.dtype), key
        st = c.state()
        if kind in ("rational", "decimator"):
            assert [st["phiIdx"], st["inputDeficit"]] == g[key + ".state"].tolist()
        if kind in ("arbitrary", "farrow"):
            assert st["inputDeficit"] == g[key + ".state"][1] and st["acc"] == g[key + ".fstate"][0]
