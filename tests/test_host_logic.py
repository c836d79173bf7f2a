"""Host logic of the library on a host-only handle (device = -1): counts, phase sequencing and carried
state must be EXACT against the oracle for every kernel type and any chunking.  CPU only."""
import ctypes as C
import math
from fractions import Fraction

import numpy as np
import pytest

import multirate_b200 as mr
import multirate_oracle as mo

F = mr._ffi


def advance(f, n):
    N = C.c_int64()
    F.check(F.lib().mrb_advance(f._handle, n, C.byref(N)))
    return N.value


def state_tuple(f):
    s = f._get_state()
    return s.phi_idx, s.input_deficit, s.phi_accumulator, s.alpha


@pytest.mark.parametrize("L,M", [(1, 1), (1, 2), (1, 8), (1, 31), (2, 1), (4, 1), (32, 1), (3, 17), (17, 3),
                                 (147, 160), (160, 147), (7, 5), (31, 32), (32, 31)])
def test_integer_sequencing_exact(L, M, rng):
    h = rng.random(int(rng.integers(1, 200)))
    f = mr.FIRFilter(h, Fraction(L, M), nchannels=1, sample_dtype=np.float64, device=-1)
    o = mo.FIRFilter(h, Fraction(L, M))
    assert type(f.kernel).__name__ == type(o.kernel).__name__
    assert f.historyLen == o.historyLen
    chunks = [0, 1, 1, 1, 2, 5, 18, 77, 0, 1, 300, 1, 1, 64, 1000]
    for n in chunks:
        assert f.outputlength(n) == o.outputlength(n) or n < o.state().get("inputDeficit", 1)
        want = len(o.filt(rng.random(n)))
        cnt = C.c_int64(); F.check(F.lib().mrb_output_count(f._handle, n, C.byref(cnt)))
        assert cnt.value == want
        assert advance(f, n) == want
        so = o.state()
        p, d, _, _ = state_tuple(f)
        assert p == so.get("phiIdx", 1) and d == so.get("inputDeficit", 1)


@pytest.mark.parametrize("polyorder", [None, 4])
@pytest.mark.parametrize("rate", [0.918734, 1.0, 1 / 2.123456789, float(np.pi), 31.7])
def test_table_sequencing_bit_exact(rate, polyorder, rng):
    """FIRArbitrary / FIRFarrow: the Float64 accumulator, alpha and deficit after every chunk are
    bit-for-bit the oracle's (src/Filters.jl:663-673, 780-786)."""
    h = rng.random(320)
    f = mr.FIRFilter(h, rate, 32, polyorder, nchannels=1, sample_dtype=np.float64, device=-1)
    o = mo.FIRFilter(h, rate, 32, polyorder)
    for n in [0, 1, 1, 1, 3, 40, 1, 777, 5000]:
        want = len(o.filt(rng.random(n)))
        assert advance(f, n) == want
        so = o.state()
        p, d, acc, alpha = state_tuple(f)
        assert d == so["inputDeficit"] and acc == so["acc"]
        if polyorder is None:
            assert p == so["phiIdx"] and alpha == so["alpha"]
        assert f.outputlength(50) == o.outputlength(50)


def test_readme_struct_fields():
    f = mr.FIRFilter(np.r_[np.ones(3), np.zeros(6)], Fraction(3, 17))
    k = f.kernel
    assert isinstance(k, mr.FIRRational)
    assert np.array_equal(k.pfb, [[0, 0, 0], [0, 0, 0], [1, 1, 1]])
    assert (k.ratio, k.Nphi, k.tapsPerphi, k.phiIdx, k.inputDeficit) == (Fraction(3, 17), 3, 3, 1, 1)
    assert f.historyLen == 2 and np.array_equal(f.history, [0.0, 0.0])


def test_kernel_selection_and_errors():
    h = np.ones(12)
    assert isinstance(mr.FIRFilter(h).kernel, mr.FIRStandard)
    assert isinstance(mr.FIRFilter(h, Fraction(3, 3)).kernel, mr.FIRStandard)        # README.md:31-32 / SURVEY 9.6
    assert isinstance(mr.FIRFilter(h, Fraction(1, 3)).kernel, mr.FIRDecimator)
    assert isinstance(mr.FIRFilter(h, Fraction(3, 1)).kernel, mr.FIRInterpolator)
    assert isinstance(mr.FIRFilter(h, Fraction(6, 8)).kernel, mr.FIRRational)
    assert mr.FIRFilter(h, Fraction(6, 8)).kernel.ratio == Fraction(3, 4)
    assert isinstance(mr.FIRFilter(h, 0.5).kernel, mr.FIRArbitrary)
    assert mr.FIRFilter(h, 0.5).kernel.Nphi == 32
    assert isinstance(mr.FIRFilter(h, 0.5, 4, 2).kernel, mr.FIRFarrow)
    with pytest.raises(ValueError, match="rate must be greater than 0"):
        mr.FIRFilter(h, -1.0)
    d = F.Desc(); d.kind = F.KIND_AUTO; d.tap_dtype = F.F64; d.sample_dtype = F.F32; d.device = -1
    hh = np.ones(4); d.h = hh.ctypes.data; d.h_len = 4; d.interpolation = 1; d.decimation = 1; d.rate = -2.0
    d.n_phi = 32; d.poly_order = -1; d.n_channels = 1
    out = C.c_void_p()
    assert F.lib().mrb_create(C.byref(d), C.byref(out)) == F.MRB_ERR_BAD_ARGUMENT
    assert b"rate must be greater than 0" in F.lib().mrb_last_error()


def test_nextphase_taps2pfb_lengths(rng):
    for L in range(1, 9):
        for M in range(1, 9):
            r = Fraction(L, M)
            for p in range(1, r.numerator + 1):
                assert mr.nextphase(p, r) == mo.nextphase(p, r)
    for n, nphi in [(9, 4), (3528, 147), (5, 7), (2336, 32)]:
        h = rng.random(n)
        assert np.array_equal(mr.taps2pfb(h, nphi), mo.taps2pfb(h, nphi))
        h32 = h.astype(np.float32)
        assert np.array_equal(mr.taps2pfb(h32, nphi), mo.taps2pfb(h32, nphi))
    for outlen, (L, M), phi in [(10, (3, 17), 2), (918750, (147, 160), 1), (7, (1, 8), 1)]:
        assert mr.inputlength(outlen, Fraction(L, M), phi) == mo.inputlength_ratio(outlen, Fraction(L, M), phi)
        assert mr.outputlength(outlen, Fraction(L, M), phi) == mo.outputlength_ratio(outlen, Fraction(L, M), phi)


def test_tapsforphase_and_banks(rng):
    h = rng.random(300).astype(np.float32)
    fa, oa = mr.FIRFilter(h, 1.3), mo.FIRFilter(h, 1.3)
    ff, of = mr.FIRFilter(h, 1.3, 32, 4), mo.FIRFilter(h, 1.3, 32, 4)
    assert np.array_equal(fa.kernel.pfb, oa.kernel.pfb) and np.array_equal(fa.kernel.dpfb, oa.kernel.dpfb)
    assert np.array_equal(ff.kernel.pnfb, of.kernel.pnfb)
    for ph in [1.0, 1.5, 7.25, 32.0, 32.999]:
        assert np.array_equal(mr.tapsforphase(fa.kernel, ph), mo.tapsforphase_arbitrary(oa.kernel, ph))
        assert np.array_equal(mr.tapsforphase(ff.kernel, ph), mo.tapsforphase_farrow(of.kernel, ph))
    assert np.array_equal(ff.kernel.currentTaps, of.kernel.currentTaps)
    with pytest.raises(ValueError, match="phase must be"):
        mr.tapsforphase(fa.kernel, 34.0)
    with pytest.raises(ValueError, match="buffer is too small"):
        mr.tapsforphase_(np.empty(2, np.float32), fa.kernel, 1.0)


def test_setphase_and_deficit_poke():
    h = np.ones(64)
    f = mr.FIRFilter(h, float(np.pi), 32, 4)
    f.kernel.inputDeficit += 3                       # examples/FIRFarrow.jl:29
    assert f.kernel.inputDeficit == 4
    assert f.setphase(0.5) == 0.5 * 31 + 1           # src/Filters.jl:226
    a = mr.FIRFilter(h, 0.7)
    phi, alpha = a.setphase(0.26)
    assert (phi, alpha) == (9, 1 + 0.26 * 32 - 9) and a.kernel.phiAccumulator == 1 + 0.26 * 32
    r = mr.FIRFilter(h, Fraction(3, 4))
    assert r.setphase(0.5) == 2 and r.setphase(1.0) == 3 and r.setphase(0.0) == 1
    with pytest.raises(AssertionError):
        r.setphase(1.5)
    with pytest.raises(mr.MrbError):
        mr.FIRFilter(h, Fraction(3, 1)).setphase(0.5)


@pytest.mark.parametrize("L,M", [(147, 160), (3, 17), (1, 8), (4, 1), (1, 1), (17, 3)])
def test_seek_closed_form(L, M, rng):
    """Segment start state: seek(n0) == the state after filtering n0 samples."""
    h = rng.random(97)
    for n0 in [0, 1, 2, 159, 160, 161, 1000, 4097, 2 ** 31 - 5]:
        a = mr.FIRFilter(h, Fraction(L, M), nchannels=1, sample_dtype=np.float32, device=-1)
        b = mr.FIRFilter(h, Fraction(L, M), nchannels=1, sample_dtype=np.float32, device=-1)
        k0 = C.c_int64()
        F.check(F.lib().mrb_seek(a._handle, n0, None, 0, C.byref(k0), None))
        assert advance(b, n0) == k0.value
        assert state_tuple(a)[:2] == state_tuple(b)[:2]


@pytest.mark.parametrize("polyorder", [None, 4])
@pytest.mark.parametrize("rate", [0.918734, 1.37, 1 / 2.123456789, 3.0001])
def test_seek_table_kinds_is_the_exact_replay(rate, polyorder, rng):
    """seek(n0) of an arbitrary / Farrow filter == the state after consuming n0 samples, bit for bit
    (phase accumulator, alpha, deficit), however the n0 samples were chunked; k0 == the outputs produced so far."""
    h = rng.random(32 * 12)
    args = (h, rate, 32) if polyorder is None else (h, rate, 32, polyorder)
    for n0 in [0, 1, 2, 31, 1000, 65536 + 17, 1_000_003]:
        a = mr.FIRFilter(*args, nchannels=1, sample_dtype=np.float32, device=-1)
        b = mr.FIRFilter(*args, nchannels=1, sample_dtype=np.float32, device=-1)
        assert a.seek(n0) == advance(b, n0 // 3) + advance(b, n0 - n0 // 3)
        assert state_tuple(a) == state_tuple(b)
        # and against the oracle's literal loop for the sizes it finishes quickly
        if n0 <= 70000:
            o = mo.FIRFilter(*args)
            assert len(o.filt(np.zeros(n0))) == a.seek(n0)
            so = o.state()
            assert a._get_state().input_deficit == so["inputDeficit"]
            if "acc" in so:
                assert a._get_state().phi_accumulator == so["acc"]


def test_set_taps_rebuilds_banks_and_keeps_state(rng):
    """SURVEY 8f rank 3 (host side): mrb_set_taps rebuilds pfb / dpfb / flipped h exactly as construction does and
    leaves the carried state alone; wrong length is refused."""
    for args in ((Fraction(3, 17),), (Fraction(1, 4),), (0.77, 8), (0.77, 8, 3)):
        h1, h2 = rng.random(50), rng.random(50)
        f = mr.FIRFilter(h1, *args, nchannels=1, sample_dtype=np.float64, device=-1)
        advance(f, 1234)
        before = state_tuple(f)
        f.set_taps(h2)
        g = mr.FIRFilter(h2, *args, nchannels=1, sample_dtype=np.float64, device=-1)
        assert np.array_equal(f._pfb(0), g._pfb(0))
        if len(args) == 2:
            assert np.array_equal(f._pfb(1), g._pfb(1))
        assert state_tuple(f) == before
        with pytest.raises(ValueError):
            f.set_taps(h2[:-1])
        with pytest.raises(mr.MrbError):
            F.check(F.lib().mrb_set_taps(f._handle, h2.ctypes.data, 49, None))


def product_schedule(f, n_in):
    N = C.c_int64()
    F.check(F.lib().mrb_output_count(f._handle, n_in, C.byref(N)))
    n, b, a = np.empty(N.value, np.int64), np.empty(N.value, np.int32), np.empty(N.value, np.float64)
    F.check(F.lib().mrb_get_schedule(f._handle, n_in, n.ctypes.data, b.ctypes.data, a.ctypes.data))
    return n, b, a


@pytest.mark.parametrize("args", [(Fraction(147, 160),), (Fraction(3, 17),), (Fraction(17, 3),), (Fraction(1, 8),),
                                  (0.918734, 32), (1.37, 32), (1 / 2.123456789, 16), (0.918734, 32, 4), (3.3, 8, 2)])
def test_schedule_bit_exact_against_the_oracle_loops(args, rng):
    """PHASE sequencing, output by output: mrb_get_schedule (what the kernels are fed) == the literal loops of the
    oracle (src/Filters.jl:558-569, 613-625, 717-732, 814-826) -- window positions, branches and the Float64 alpha /
    phase bit for bit, over ragged chunks with carried state, a setphase in between for the kinds that have one."""
    h = rng.random(32 * 9)
    f = mr.FIRFilter(h, *args, nchannels=1, sample_dtype=np.float64, device=-1)
    o = mo.FIRFilter(h, *args)
    for i, n in enumerate([1, 2, 700, 3, 4096, 0, 1531, 65536]):
        if i == 4 and not (len(args) == 1 and args[0].numerator == 1):
            assert f.setphase(0.37) == o.setphase(0.37)
        pn, pb, pa = product_schedule(f, n)
        advance(f, n)
        o.last_schedule = ([], [], [])
        y = o.filt(np.zeros(n))
        sched = o.last_schedule
        assert len(pn) == len(y) == len(sched[0])
        assert np.array_equal(pn, np.asarray(sched[0], dtype=np.int64) - 1)
        if len(args) == 3:                                       # farrow: Float64 phase
            assert np.array_equal(pa, np.asarray(sched[1], dtype=np.float64))
        else:
            assert np.array_equal(pb, np.asarray(sched[1], dtype=np.int64) - 1)
            if len(args) == 2:                                   # arbitrary: alpha
                assert np.array_equal(pa, np.asarray(sched[2], dtype=np.float64))


def _literal_replay(acc, x_idx, delta, nphi, n_in):
    """The reference's update (src/Filters.jl:663-673), literally, in Python floats: (n, acc) per output and the end state."""
    ns, accs = [], []
    while x_idx <= n_in:
        ns.append(x_idx - 1)
        accs.append(acc)
        acc += delta
        if acc > nphi:
            x_idx += int(math.floor((acc - 1) / nphi))
            acc = math.fmod(acc - 1, nphi) + 1
    return np.asarray(ns, np.int64), np.asarray(accs, np.float64), acc, x_idx


def test_replay_equals_the_literal_recurrence_for_random_rates(rng):
    """mrb_seq.h ArbStepper (one add and one exact subtract per update, a guard band for the rounded quotient) against the
    literal floating-point update of the reference: random rates from steep decimation to high interpolation, random branch
    counts and start accumulators set by the caller, output by output and bit for bit."""
    h = rng.random(64)
    cases = [(0.918734, 32), (1.37, 32), (0.5000001, 32), (0.51, 7), (15.9, 32), (1.0, 32), (2.0, 32), (1 / 3.0, 5), (40.0, 32)]
    for _ in range(40):
        nphi = int(rng.integers(1, 70))
        cases.append((float(np.exp(rng.uniform(np.log(0.3), np.log(nphi + 3.0)))), nphi))
    for rate, nphi in cases:
        f = mr.FIRFilter(h, rate, nphi, nchannels=1, sample_dtype=np.float64, device=-1)
        delta = nphi / rate
        acc, x_idx = 1.0, 1
        for j, n_in in enumerate([3, 5000, 1, 12000, 64, 9000]):
            if j == 3:                                              # an accumulator the caller set: not on any grid
                s = f._get_state()
                acc = float(rng.uniform(1.0, nphi + 1.0))
                s.phi_accumulator = acc
                s.phi_idx, s.alpha = int(acc), acc - int(acc)
                f._set_state(s)
                x_idx = s.input_deficit
            pn, pb, pa = product_schedule(f, n_in)
            wn, wacc, acc, x_idx = _literal_replay(acc, x_idx, delta, float(nphi), n_in)
            assert len(pn) == len(wn), (rate, nphi, j)
            assert np.array_equal(pn, wn), (rate, nphi, j)
            if len(wn) > 1:                                         # (output 0 carries the stored (phi, alpha) pair)
                wphi = np.floor(wacc[1:]).astype(np.int64)
                assert np.array_equal(pb[1:], wphi - 1), (rate, nphi, j)
                assert np.array_equal(pa[1:], wacc[1:] - wphi), (rate, nphi, j)
            assert advance(f, n_in) == len(wn)
            s = f._get_state()
            x_idx -= n_in                                           # the carried deficit (src/Filters.jl:734)
            assert s.phi_accumulator == acc and s.input_deficit == x_idx, (rate, nphi, j, s.phi_accumulator, acc)
