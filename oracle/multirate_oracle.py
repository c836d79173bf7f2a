"""CPU restatement of Multirate.jl's streaming polyphase FIR path (numpy).

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`multirate.jl_b200/`) may
import this file; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it, and only as the
checker.

Every function follows one place in the reference (`/root/reference`, cited as
file:line) as a *literal sequential state machine*: the per-output index /
phase bookkeeping is replayed step by step exactly as the Julia loops do it
(1-based indices and all), so output COUNTS, PHASE sequences and carried STATE
are exact.  Output VALUES are accumulated in a wider type than the reference
uses (float64 for Float32 paths, long double for Float64 paths) because the
reference's `@simd` loops (src/support.jl:9,23,26,37,47,50) do not pin a
summation order; value parity is therefore tolerance based (1e-5 / 1e-12,
normalised by max|y|), count/phase/state parity is exact.

Parity pinning: checked in tests/test_oracle.py against every known-answer
vector the reference holds for this path -- README.md:58-142 (3//17 example,
values + printed pfb + initial state), src/Filters.jl:276-280 (taps2pfb
example), test/runtests.jl:423-438 (nextphase), the notebook's cell 10 (Farrow
rate pi, 40 in -> 126 out) -- and against the textbook definition the
reference's own tests use (zero-stuff, lfilter, stride: test/runtests.jl:
123-124,190-194,270-278).  FIRFarrow VALUES are unpinned by the reference
(test/farrowtest.jl only prints); they are pinned here by this restatement with
the conventions stated at `polyfit` / `polyval` below: "parity unpinned" for
Farrow values, pinned for everything else.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

_REAL_OF = {np.dtype(np.float32): np.float32, np.dtype(np.float64): np.float64,
            np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}


def promote_type(th, tx):
    """Julia promote_type(Th, Tx) for Th in {f32,f64}, Tx in {f32,f64,c64,c128}."""
    return np.result_type(np.dtype(th), np.dtype(tx))


def _wide(dtype):
    """Accumulation type used by the oracle for a given output dtype."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return np.float64
    if dtype == np.complex64:
        return np.complex128
    if dtype == np.float64:
        return np.longdouble
    return np.clongdouble


# --------------------------------------------------------------------------
# bank construction
# --------------------------------------------------------------------------
def taps2pfb(h, Nphi):
    """src/Filters.jl:284-298.  pfb[T-1-r, c] = h[r*Nphi + c] (0-based), zero padded:
    column phi is polyphase branch phi, time reversed."""
    h = np.asarray(h)
    hLen = len(h)
    T = -(-hLen // Nphi)
    pfb = np.zeros((T, Nphi), dtype=h.dtype)
    hIdx = 0
    for rowIdx in range(T - 1, -1, -1):
        for colIdx in range(Nphi):
            pfb[rowIdx, colIdx] = h[hIdx] if hIdx < hLen else 0
            hIdx += 1
    return pfb


def polyfit(y, polyorder):
    """src/support.jl:85-88: least squares A\\y, A[x,p] = x^p, x = 1..len(y),
    p = 0..order; coefficients lowest order first (Polynomials.Poly).
    Convention fixed here (unpinned upstream): the solve is done in Float64."""
    y = np.asarray(y, dtype=np.float64)
    xs = np.arange(1, len(y) + 1, dtype=np.float64)
    A = np.vander(xs, polyorder + 1, increasing=True)
    coef, *_ = np.linalg.lstsq(A, y, rcond=None)
    return coef


# The Farrow coefficients are an INPUT of the filtering path (fitted once, at construction).  The least-squares problem
# is ill-conditioned (cond ~2.4e6 at order 4), so two correct solvers differ around the 10th digit -- more than the
# Float64 filtering tolerance (1e-12).  Filtering parity therefore takes the coefficients AS DATA: a test may install
# the fit it wants both sides to use (tests/conftest.py installs the library's mrb_pfb2pnfb); the solve below stays the
# oracle's own, independent one (numpy SVD) and tests/test_farrow_fit.py compares the two.
_PNFB_PROVIDER = None


def set_pnfb_provider(fn):
    """fn(pfb, polyorder) -> (T, order+1) float64 coefficients, or None to restore the oracle's own fit."""
    global _PNFB_PROVIDER
    _PNFB_PROVIDER = fn


def pfb2pnfb(pfb, polyorder):
    """src/Filters.jl:311-321: one polynomial per tap ROW of the bank, fitted over
    phi = 1..Nphi.  Stored as Poly{T} (src/Filters.jl:313), i.e. coefficients are
    rounded to the tap type T; kept in a float64 array holding T-representable
    values.  Shape (T, order+1), lowest order first."""
    T, Nphi = pfb.shape
    out = np.empty((T, polyorder + 1), dtype=np.float64)
    for i in range(T):
        out[i] = polyfit(pfb[i, :], polyorder).astype(pfb.dtype).astype(np.float64)
    return out


def polyval(coefs, x):
    """Polynomials.polyval: Horner, highest order first, in promote(T, Float64) =
    Float64.  `coefs` (..., order+1) lowest first; x scalar float."""
    coefs = np.asarray(coefs, dtype=np.float64)
    acc = coefs[..., -1].copy()
    for p in range(coefs.shape[-1] - 2, -1, -1):
        acc = acc * np.float64(x) + coefs[..., p]
    return acc


def nextphase(currentphase, ratio):
    """src/Filters.jl:433-439 (1-based phases)."""
    ratio = Fraction(ratio)
    interpolation, decimation = ratio.numerator, ratio.denominator
    step = decimation % interpolation
    nxt = currentphase + step
    return nxt - interpolation if nxt > interpolation else nxt


def outputlength_ratio(inputlength, ratio, initialphi):
    """src/Filters.jl:352-357 (Float64 division then ceil)."""
    ratio = Fraction(ratio)
    return int(math.ceil(((inputlength * ratio.numerator) - initialphi + 1) / ratio.denominator))


def inputlength_ratio(outputlength, ratio, initialphi):
    """src/Filters.jl:396-401."""
    ratio = Fraction(ratio)
    return int(math.ceil((outputlength * ratio.denominator + initialphi - 1) / ratio.numerator))


def shiftin(a, b):
    """src/support.jl:61-80: a = [a; b][end-len(a)+1:end] (last axis)."""
    aLen = a.shape[-1]
    if aLen == 0:
        return a
    return np.concatenate([a, b], axis=-1)[..., -aLen:].copy()


# --------------------------------------------------------------------------
# kernels (state carriers) -- src/Filters.jl:15-147
# --------------------------------------------------------------------------
class FIRStandard:
    def __init__(self, h):
        self.h = np.asarray(h)[::-1].copy()            # flipud, :21
        self.hLen = len(h)


class FIRInterpolator:
    def __init__(self, h, interpolation):
        self.pfb = taps2pfb(h, interpolation)           # :36
        self.interpolation = interpolation
        self.tapsPerphi, self.Nphi = self.pfb.shape


class FIRDecimator:
    def __init__(self, h, decimation):
        self.h = np.asarray(h)[::-1].copy()             # :53
        self.hLen = len(h)
        self.decimation = decimation
        self.inputDeficit = 1                            # :56


class FIRRational:
    def __init__(self, h, ratio):
        ratio = Fraction(ratio)
        self.pfb = taps2pfb(h, ratio.numerator)         # :73
        self.ratio = ratio
        self.tapsPerphi, self.Nphi = self.pfb.shape
        self.criticalYidx = int(math.floor(self.tapsPerphi * ratio))
        self.phiIdx = 1                                  # :77
        self.inputDeficit = 1                            # :78


class FIRArbitrary:
    def __init__(self, h, rate, Nphi):
        h = np.asarray(h)
        dh = np.concatenate([np.diff(h), np.zeros(1, h.dtype)]).astype(h.dtype)   # :106
        self.rate = float(rate)
        self.pfb = taps2pfb(h, Nphi)                    # :107
        self.dpfb = taps2pfb(dh, Nphi)                  # :108
        self.Nphi = Nphi
        self.tapsPerphi = self.pfb.shape[0]
        self.phiAccumulator = 1.0
        self.phiIdx = 1
        self.alpha = 0.0
        self.delta = Nphi / self.rate                    # :113
        self.inputDeficit = 1
        self.xIdx = 1

    def update(self):
        """src/Filters.jl:663-673."""
        self.phiAccumulator += self.delta
        if self.phiAccumulator > self.Nphi:
            self.xIdx += int(math.floor((self.phiAccumulator - 1) / self.Nphi))
            self.phiAccumulator = math.fmod(self.phiAccumulator - 1, self.Nphi) + 1
        self.phiIdx = int(math.floor(self.phiAccumulator))
        self.alpha = self.phiAccumulator - self.phiIdx


class FIRFarrow:
    def __init__(self, h, rate, Nphi, polyorder, pnfb=None):
        h = np.asarray(h)
        self.rate = float(rate)
        self.pfb = taps2pfb(h, Nphi)                    # :138
        if pnfb is not None:                            # coefficients as data
            self.pnfb = np.asarray(pnfb, dtype=np.float64).reshape(self.pfb.shape[0], polyorder + 1)
        elif _PNFB_PROVIDER is not None:
            self.pnfb = np.asarray(_PNFB_PROVIDER(self.pfb, polyorder), dtype=np.float64)
        else:
            self.pnfb = pfb2pnfb(self.pfb, polyorder)   # :139
        self.polyorder = polyorder
        self.Nphi = Nphi
        self.tapsPerphi = self.pfb.shape[0]
        self.phiIdx = 1.0
        self.delta = Nphi / self.rate
        self.inputDeficit = 1
        self.xIdx = 1
        self.currentTaps = polyval(self.pnfb, self.phiIdx).astype(h.dtype)   # :145

    def update(self):
        """src/Filters.jl:780-792 (taps are evaluated lazily per output by filt)."""
        self.phiIdx += self.delta
        if self.phiIdx > self.Nphi:
            self.xIdx += int(math.floor((self.phiIdx - 1) / self.Nphi))
            self.phiIdx = math.fmod(self.phiIdx - 1, self.Nphi) + 1


# --------------------------------------------------------------------------
# FIRFilter -- src/Filters.jl:151-198
# --------------------------------------------------------------------------
class FIRFilter:
    """FIRFilter(h, ratio::Fraction|int = 1) / FIRFilter(h, rate::float, Nphi=32) /
    FIRFilter(h, rate::float, Nphi, polyorder).

    `nchannels` is the one addition: `filt` then takes x of shape (nchannels, n)
    and applies the SAME state machine to every row (the reference has one
    FIRFilter per vector; channels never interact).  The sample dtype is fixed by
    the first `filt` call (src/Filters.jl:452: history is converted to Vector{Tx})."""

    def __init__(self, h, ratio=Fraction(1, 1), Nphi=None, polyorder=None, pnfb=None):
        h = np.asarray(h)
        assert h.dtype in (np.float32, np.float64)
        if isinstance(ratio, float):
            if not ratio > 0.0:
                raise ValueError("rate must be greater than 0")          # :184,193
            if polyorder is None:
                self.kernel = FIRArbitrary(h, ratio, 32 if Nphi is None else Nphi)   # :183-189
            else:
                self.kernel = FIRFarrow(h, ratio, Nphi, polyorder, pnfb)  # :192-198
            self.historyLen = self.kernel.tapsPerphi - 1
        else:
            ratio = Fraction(ratio)
            interpolation, decimation = ratio.numerator, ratio.denominator
            if ratio == 1:
                self.kernel = FIRStandard(h)
                self.historyLen = self.kernel.hLen - 1                   # :165
            elif interpolation == 1:
                self.kernel = FIRDecimator(h, decimation)
                self.historyLen = self.kernel.hLen - 1                   # :168
            elif decimation == 1:
                self.kernel = FIRInterpolator(h, interpolation)
                self.historyLen = self.kernel.tapsPerphi - 1             # :171
            else:
                self.kernel = FIRRational(h, ratio)
                self.historyLen = self.kernel.tapsPerphi - 1             # :174
        self.th = h.dtype
        self.history = None          # allocated at first filt (dtype Tx, per channel)

    # ---- helpers -----------------------------------------------------
    def _prep(self, x):
        x = np.asarray(x)
        self._squeeze = x.ndim == 1
        x2 = x[None, :] if x.ndim == 1 else x
        if self.history is None or self.history.shape[0] != x2.shape[0] or self.history.dtype != x2.dtype:
            self.history = np.zeros((x2.shape[0], self.historyLen), dtype=x2.dtype)    # :177 + :452
        return x2

    def _dots(self, taps_cols, x2, n_idx):
        """y[c,k] = sum_i taps_cols[i,k] * ext[c, n_idx[k]-T+i]  with ext = [history | x],
        n_idx 1-based index into x of the LAST window sample (may be < T: straddles
        history: src/support.jl:16-31,44-55 vs :5-14,33-42)."""
        out_dtype = promote_type(self.th, x2.dtype)
        wide = _wide(out_dtype)
        T = taps_cols.shape[0]
        K = len(n_idx)
        y = np.zeros((x2.shape[0], K), dtype=out_dtype)
        if K == 0:
            return y
        ext = np.concatenate([self.history, x2], axis=1)
        H = self.historyLen                      # == T-1 for every kernel
        base = np.asarray(n_idx, dtype=np.int64) - 1 + H - (T - 1)     # 0-based start in ext
        idx = base[:, None] + np.arange(T)[None, :]                    # (K, T)
        tw = taps_cols.T.astype(wide)                                  # (K, T)
        for c in range(x2.shape[0]):
            y[c] = (ext[c][idx].astype(wide) * tw).sum(axis=1).astype(out_dtype)
        return y

    def _finish(self, y):
        return y[0] if self._squeeze else y

    # ---- filt ---------------------------------------------------------
    def filt(self, x):
        k = self.kernel
        x2 = self._prep(x)
        xLen = x2.shape[1]
        out_dtype = promote_type(self.th, x2.dtype)
        empty = np.zeros((x2.shape[0], 0), dtype=out_dtype)

        if isinstance(k, FIRStandard):                                   # :450-473
            n_idx = list(range(1, xLen + 1))
            y = self._dots(np.repeat(k.h[:, None], xLen, axis=1), x2, n_idx)
            self.history = shiftin(self.history, x2)
            return self._finish(y)

        if isinstance(k, FIRInterpolator):                               # :489-517
            n_idx, phis = [], []
            phi, inputIdx = 1, 1
            for _ in range(k.interpolation * xLen):
                n_idx.append(inputIdx); phis.append(phi)
                if phi == k.Nphi:
                    phi, inputIdx = 1, inputIdx + 1
                else:
                    phi += 1
            y = self._dots(k.pfb[:, np.asarray(phis, dtype=np.int64) - 1], x2, n_idx)
            self.history = shiftin(self.history, x2)
            return self._finish(y)

        if isinstance(k, FIRDecimator):                                  # :598-650 (+ SURVEY 9.3)
            if xLen < k.inputDeficit:
                self.history = shiftin(self.history, x2)
                k.inputDeficit -= xLen
                return self._finish(empty)
            n_idx = []
            inputIdx = k.inputDeficit
            while inputIdx <= xLen:
                n_idx.append(inputIdx)
                inputIdx += k.decimation
            k.inputDeficit = inputIdx - xLen
            self.last_schedule = (n_idx, [1] * len(n_idx))
            y = self._dots(np.repeat(k.h[:, None], len(n_idx), axis=1), x2, n_idx)
            self.history = shiftin(self.history, x2)
            return self._finish(y)

        if isinstance(k, FIRRational):                                   # :536-575
            if xLen < k.inputDeficit:
                self.history = shiftin(self.history, x2)
                k.inputDeficit -= xLen
                return self._finish(empty)
            interpolation, decimation = k.ratio.numerator, k.ratio.denominator
            n_idx, phis = [], []
            inputIdx = k.inputDeficit
            while inputIdx <= xLen:
                n_idx.append(inputIdx); phis.append(k.phiIdx)
                inputIdx += int(math.floor((k.phiIdx + decimation - 1) / interpolation))   # :567
                k.phiIdx = nextphase(k.phiIdx, k.ratio)                                    # :568
            k.inputDeficit = inputIdx - xLen
            self.last_schedule = (n_idx, phis)
            y = self._dots(k.pfb[:, np.asarray(phis, dtype=np.int64) - 1], x2, n_idx)
            self.history = shiftin(self.history, x2)
            return self._finish(y)

        if isinstance(k, FIRArbitrary):                                  # :693-742
            if xLen < k.inputDeficit:
                self.history = shiftin(self.history, x2)
                k.inputDeficit -= xLen
                return self._finish(empty)
            k.xIdx = k.inputDeficit
            n_idx, phis, alphas = [], [], []
            while k.xIdx <= xLen:
                n_idx.append(k.xIdx); phis.append(k.phiIdx); alphas.append(k.alpha)
                k.update()
            k.inputDeficit = k.xIdx - xLen
            self.last_schedule = (n_idx, phis, alphas)
            ph = np.asarray(phis, dtype=np.int64) - 1
            wide = _wide(out_dtype)
            # yLower + yUpper*alpha (:730); the two dots are kept separate as upstream.
            yl = self._dots_wide(k.pfb[:, ph], x2, n_idx, wide)
            yu = self._dots_wide(k.dpfb[:, ph], x2, n_idx, wide)
            y = (yl + yu * np.asarray(alphas, dtype=np.float64)[None, :]).astype(out_dtype)
            self.history = shiftin(self.history, x2)
            return self._finish(y)

        if isinstance(k, FIRFarrow):                                     # :795-836
            if xLen < k.inputDeficit:
                self.history = shiftin(self.history, x2)
                k.inputDeficit -= xLen
                return self._finish(empty)
            k.xIdx = k.inputDeficit
            n_idx, phis = [], []
            while k.xIdx <= xLen:
                n_idx.append(k.xIdx); phis.append(k.phiIdx)
                k.update()
            k.inputDeficit = k.xIdx - xLen
            self.last_schedule = (n_idx, phis)
            # currentTaps[i] = polyval(pnfb[i], phiIdx) rounded to T (:789-791)
            taps = np.empty((k.tapsPerphi, len(phis)), dtype=self.th)
            for j, p in enumerate(phis):
                taps[:, j] = polyval(k.pnfb, p).astype(self.th)
            k.currentTaps = polyval(k.pnfb, k.phiIdx).astype(self.th)
            y = self._dots(taps, x2, n_idx)
            self.history = shiftin(self.history, x2)
            return self._finish(y)

        raise TypeError(type(k))

    def _dots_wide(self, taps_cols, x2, n_idx, wide):
        T = taps_cols.shape[0]
        K = len(n_idx)
        y = np.zeros((x2.shape[0], K), dtype=wide)
        if K == 0:
            return y
        ext = np.concatenate([self.history, x2], axis=1)
        base = np.asarray(n_idx, dtype=np.int64) - 1 + self.historyLen - (T - 1)
        idx = base[:, None] + np.arange(T)[None, :]
        tw = taps_cols.T.astype(wide)
        for c in range(x2.shape[0]):
            y[c] = (ext[c][idx].astype(wide) * tw).sum(axis=1)
        return y

    # ---- outputlength (src/Filters.jl:359-385) -------------------------
    def outputlength(self, inputlength):
        k = self.kernel
        if isinstance(k, FIRStandard):
            return inputlength
        if isinstance(k, FIRInterpolator):
            return k.interpolation * inputlength
        if isinstance(k, FIRDecimator):
            return outputlength_ratio(inputlength - k.inputDeficit + 1, Fraction(1, k.decimation), 1)
        if isinstance(k, FIRRational):
            return outputlength_ratio(inputlength - k.inputDeficit + 1, k.ratio, k.phiIdx)
        return int(math.ceil((inputlength - k.inputDeficit + 1) * k.rate))

    # ---- reset (full re-initialisation; superset of :244-260) ----
    def reset(self):
        k = self.kernel
        if self.history is not None:
            self.history = np.zeros_like(self.history)
        if isinstance(k, (FIRDecimator, FIRRational, FIRArbitrary, FIRFarrow)):
            k.inputDeficit = 1
        if isinstance(k, FIRRational):
            k.phiIdx = 1
        if isinstance(k, FIRArbitrary):
            k.phiAccumulator, k.phiIdx, k.alpha, k.xIdx = 1.0, 1, 0.0, 1
        if isinstance(k, FIRFarrow):
            k.phiIdx, k.xIdx = 1.0, 1
            k.currentTaps = polyval(k.pnfb, k.phiIdx).astype(self.th)
        return self

    def setphase(self, phi):
        """src/Filters.jl:210-232 with the repairs SURVEY section 9 decided (items 1 and 8): the Rational method reads
        an undefined variable upstream (here phi in [0,1] maps onto branch 1..Nphi); the Arbitrary method upstream can
        yield branch 0 and leaves the accumulator stale (here the accumulator is set and branch / alpha derive from it);
        the Farrow method is upstream's (:224-229)."""
        if not (0 <= phi <= 1):
            raise AssertionError("phase must be in [0, 1]")
        k = self.kernel
        if isinstance(k, FIRRational):
            k.phiIdx = min(int(math.floor(phi * k.Nphi)) + 1, k.Nphi)
            return k.phiIdx
        if isinstance(k, FIRArbitrary):
            k.phiAccumulator = 1.0 + phi * k.Nphi
            k.phiIdx = min(int(math.floor(k.phiAccumulator)), k.Nphi)
            k.alpha = k.phiAccumulator - k.phiIdx
            return k.phiIdx, k.alpha
        if isinstance(k, FIRFarrow):
            k.phiIdx = phi * (k.Nphi - 1) + 1                                       # :226
            k.currentTaps = tapsforphase_farrow(k, k.phiIdx)                        # :227
            return k.phiIdx
        raise ValueError("setphase is not supported for this kernel (it carries no phase)")

    def state(self):
        """(phase index 1-based, inputDeficit, accumulator, alpha) for state read-back tests."""
        k = self.kernel
        if isinstance(k, FIRRational):
            return dict(phiIdx=k.phiIdx, inputDeficit=k.inputDeficit)
        if isinstance(k, FIRDecimator):
            return dict(inputDeficit=k.inputDeficit)
        if isinstance(k, FIRArbitrary):
            return dict(phiIdx=k.phiIdx, inputDeficit=k.inputDeficit, acc=k.phiAccumulator, alpha=k.alpha)
        if isinstance(k, FIRFarrow):
            return dict(inputDeficit=k.inputDeficit, acc=k.phiIdx)
        return {}


def filt(h, x, ratio=Fraction(1, 1), Nphi=None, polyorder=None):
    """One-shot forms, src/Filters.jl:858-873."""
    return FIRFilter(h, ratio, Nphi, polyorder).filt(x)


def tapsforphase_arbitrary(kernel, phase):
    """src/Filters.jl:677-688."""
    if not (0 <= phase <= kernel.Nphi + 1):
        raise ValueError("phase must be >= 0 and <= Nphi+1")
    alpha, phiIdx = math.modf(phase)
    phiIdx = int(phiIdx)
    return (kernel.pfb[:, phiIdx - 1].astype(np.float64) + alpha * kernel.dpfb[:, phiIdx - 1].astype(np.float64)).astype(kernel.pfb.dtype)


def tapsforphase_farrow(kernel, phase):
    """src/Filters.jl:764-773."""
    if not (0 <= phase <= kernel.Nphi + 1):
        raise ValueError("phase must be >= 0 and <= Nphi+1")
    return polyval(kernel.pnfb, phase).astype(kernel.pfb.dtype)


# --------------------------------------------------------------------------
# second, independent oracle: the textbook definition used by the reference's
# own tests (test/runtests.jl:123-124,190-194,270-278; src/NaiveResamplers.jl:5-18)
# --------------------------------------------------------------------------
def naivefilt(h, x, ratio=Fraction(1, 1)):
    from scipy.signal import lfilter
    ratio = Fraction(ratio)
    up, down = ratio.numerator, ratio.denominator
    x = np.asarray(x)
    wide = _wide(promote_type(np.asarray(h).dtype, x.dtype))
    if wide in (np.longdouble, np.clongdouble):      # lfilter has no long double path
        wide = np.float64 if wide == np.longdouble else np.complex128
    xs = np.zeros(len(x) * up, dtype=wide)
    xs[::up] = x
    y = lfilter(np.asarray(h, dtype=np.float64), 1.0, xs)
    return y[::down]


def naivefilt_arbitrary(h, x, rate, numfilters=32):
    """src/NaiveResamplers.jl:24-49 -- loose (~1e-4) sanity oracle only: FIRArbitrary
    differs from it by construction."""
    xi = naivefilt(h, x, Fraction(numfilters, 1))
    xLen = len(xi)
    y = []
    xIdx, alpha = 1, 0.0
    delta, stride = math.modf(numfilters / rate)
    stride = int(stride)
    while xIdx < xLen:
        lo, up = xi[xIdx - 1], xi[xIdx]
        y.append(lo + alpha * (up - lo))
        alpha += delta
        xIdx += int(math.floor(alpha)) + stride
        alpha = math.fmod(alpha, 1.0)
    return np.asarray(y)


# --------------------------------------------------------------------------
# tap design twin (src/FIRDesign.jl:18-95) -- only to GENERATE benchmark/test taps
# --------------------------------------------------------------------------
def kaiserlength(transition, attenuation=60, samplerate=1.0):
    transition = transition / samplerate
    numtaps = int(math.ceil((attenuation - 7.95) / (2 * math.pi * 2.285 * transition)))
    if attenuation > 50:
        beta = 0.1102 * (attenuation - 8.7)
    elif attenuation >= 21:
        beta = 0.5842 * (attenuation - 21) ** 0.4 + 0.07886 * (attenuation - 21)
    else:
        beta = 0.0
    return numtaps, beta


def firdes(numtaps, cutoff, beta=6.75, samplerate=1.0):
    """Low-pass windowed sinc, src/FIRDesign.jl:52,76-86 (Kaiser window taken as
    numpy.kaiser(n, beta); taps are an INPUT to the path, so the exact window
    convention does not affect parity)."""
    F = cutoff / samplerate
    M = numtaps - 1
    n = np.arange(numtaps, dtype=np.float64)
    return 2 * F * np.sinc(2 * F * (n - M / 2)) * np.kaiser(numtaps, beta)
