"""Parity of the CUDA path (through the C-ABI, libmrb.so) against the oracle.  Needs a B200.

Bar (north_star): output COUNT, PHASE sequencing and carried STATE bit-exact for every kernel type and
chunking, including the empty-output cases; VALUES within 1e-5 (Float32 / Complex64) and 1e-12
(Float64 / Complex128) of the oracle, error normalised by max|y|."""
import ctypes as C
import json
import os
from fractions import Fraction

import numpy as np
import pytest

import multirate_b200 as mr
import multirate_oracle as mo
from conftest import nerr, rand_samples, tol_for

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "kat.json")))
F = mr._ffi


def states_equal(f, o):
    s, so = f._get_state(), o.state()
    ok = s.phi_idx == so.get("phiIdx", s.phi_idx) and s.input_deficit == so.get("inputDeficit", 1)
    if "acc" in so:
        ok = ok and s.phi_accumulator == so["acc"]
    if "alpha" in so:
        ok = ok and s.alpha == so["alpha"]
    return ok


def test_extension_is_loaded_and_device_is_b200():
    import torch
    assert torch.cuda.is_available()
    assert torch.cuda.get_device_capability(0)[0] == 10
    assert os.path.exists(F.LIB)
    maps = open("/proc/self/maps").read()
    F.lib()
    assert "libmrb.so" in open("/proc/self/maps").read() or "libmrb.so" in maps


def test_readme_3_17_kat_on_gpu():
    k = KAT["readme_3_17"]
    h, x = np.array(k["h"], dtype=np.float64), np.array(k["x"], dtype=np.float64)
    f = mr.FIRFilter(h, Fraction(*k["ratio"]))
    pos, ys = 0, []
    for n, want in zip(k["chunks"], k["y"]):
        y = mr.filt(f, x[pos:pos + n]); pos += n
        assert np.array_equal(y, np.array(want)), (y, want)
        ys.append(y)
    assert np.sum(np.concatenate(ys) - mr.filt(h, x, Fraction(*k["ratio"]))) == 0.0
    o = mo.FIRFilter(h, Fraction(*k["ratio"])); o.filt(x)
    assert states_equal(f, o)


def test_farrow_notebook_count_on_gpu():
    k = KAT["farrow_notebook_count"]
    N = k["Nphi"]
    h = mo.firdes(k["tapsPerphi"] * N, min(0.45 / N, k["rate"] / N)) * N
    t = np.arange(k["n_in"])
    x = np.cos(2 * np.pi * 0.15 * t) + 0.5 * np.sin(2 * np.pi * 0.3 * t * np.pi)
    assert len(mr.filt(h, x, k["rate"], N, k["polyorder"])) == k["n_out"]


def test_golden_oracle_vectors():
    """Frozen vectors (tests/golden/oracle_vectors.npz): all 7 cases x 2 tap dtypes x 4 sample dtypes, 3 chunks
    (1 sample / 39 / 291), 2 channels, values + end state."""
    g = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
    names = sorted({k.rsplit(".", 1)[0] for k in g.files})
    ratios = {"standard": Fraction(1, 1), "decimator": Fraction(1, 8), "interpolator": Fraction(4, 1),
              "rational": Fraction(147, 160), "rational_3_17": Fraction(3, 17)}
    for key in names:
        name, th, tx = key.split(".")
        h, x = g[key + ".h"], g[key + ".x"]
        if name == "arbitrary":
            f = mr.FIRFilter(h, 0.918734, 32)
        elif name == "farrow":
            f = mr.FIRFilter(h, 0.918734, 32, 4, pnfb=g[key + ".pnfb"])      # coefficients as data, frozen with the vectors
        else:
            f = mr.FIRFilter(h, ratios[name])
        for i, (a, b) in enumerate(((0, 1), (1, 40), (40, 331))):
            y = f.filt(x[:, a:b])
            want = g[key + ".y%d" % i]
            assert y.dtype == want.dtype and y.shape == want.shape, key
            assert nerr(y, want) <= tol_for(y.dtype), (key, i, nerr(y, want))
        s = f._get_state()
        if name.startswith("rational") or name == "decimator":
            assert [s.phi_idx, s.input_deficit] == g[key + ".state"].tolist(), key
        if name in ("arbitrary", "farrow"):
            assert s.input_deficit == g[key + ".state"][1] and s.phi_accumulator == g[key + ".fstate"][0], key
        if name == "arbitrary":
            assert s.alpha == g[key + ".fstate"][1]


@pytest.mark.parametrize("th", [np.float32, np.float64])
@pytest.mark.parametrize("tx", [np.float32, np.float64, np.complex64, np.complex128])
def test_four_way_equivalence_gpu(th, tx, rng):
    """The reference's own test matrix (test/runtests.jl:389-421), seeded: naive definition == oracle ==
    GPU one-shot == GPU 2-chunk == GPU sample-at-a-time (exercises empty returns and deficit carry)."""
    Ls = [1] + sorted(set(rng.integers(2, 33, 3).tolist()))
    Ms = [1] + sorted(set(rng.integers(2, 33, 3).tolist()))
    for L in Ls:
        for M in Ms:
            ratio = Fraction(L, M)
            h = rng.random(int(rng.integers(16, 129))).astype(th)
            xLen = int(rng.integers(200, 301)); xLen -= xLen % M
            x = rand_samples(rng, xLen, tx)
            ty = np.result_type(th, tx)
            tol = tol_for(ty)
            naive = mo.naivefilt(h, x, ratio)
            want = mo.filt(h, x, ratio)
            one = mr.filt(h, x, ratio)
            assert one.dtype == ty and one.shape == want.shape
            assert nerr(one, want) <= tol, (L, M, nerr(one, want))
            assert nerr(one, naive.astype(ty)) <= max(tol, 2e-6)
            f, o = mr.FIRFilter(h, ratio), mo.FIRFilter(h, ratio)
            piv = min(int(rng.integers(50, 151)), xLen // 4)
            two = np.concatenate([f.filt(x[:piv]), f.filt(x[piv:])])
            o.filt(x[:piv]); o.filt(x[piv:])
            assert nerr(two, want) <= tol and states_equal(f, o)
            f.reset(); o.reset()
            parts = []
            for i in range(piv):                                     # test/runtests.jl:81-83,150-152,311-313
                yi, oi = f.filt(x[i:i + 1]), o.filt(x[i:i + 1])
                assert len(yi) == len(oi)                            # empty returns included
                parts.append(yi)
            parts.append(f.filt(x[piv:])); o.filt(x[piv:])
            assert nerr(np.concatenate(parts), want) <= tol and states_equal(f, o)


@pytest.mark.parametrize("th,tx", [(np.float32, np.float32), (np.float32, np.complex64), (np.float64, np.float64),
                                   (np.float64, np.complex128), (np.float64, np.float32), (np.float32, np.float64)])
@pytest.mark.parametrize("polyorder", [None, 4])
def test_arbitrary_and_farrow_gpu(th, tx, polyorder, rng):
    """FIRArbitrary / FIRFarrow against the oracle: counts, accumulator, alpha and deficit bit-exact after every
    chunk (1-sample chunks included), values within tolerance; multi-channel."""
    N = 32
    hLen, beta = mo.kaiserlength(0.05, samplerate=N)
    hLen = -(-hLen // N) * N
    h = (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(th)          # test/runtests.jl:336-341
    x = rand_samples(rng, (3, 700), tx)
    for rate in (0.918734, 1 / 2.123456789, 2.5):
        f, o = mr.FIRFilter(h, rate, N, polyorder), mo.FIRFilter(h, rate, N, polyorder)
        pos = 0
        for n in [1, 1, 0, 3, 95, 1, 600 - 101]:
            y, w = f.filt(x[:, pos:pos + n]), o.filt(x[:, pos:pos + n]); pos += n
            assert y.shape == w.shape and y.dtype == w.dtype
            assert nerr(y, w) <= tol_for(w.dtype), (rate, n, nerr(y, w))
            assert states_equal(f, o)
        one = mr.filt(h, x[0], rate, N, polyorder) if polyorder is not None else mr.filt(h, x[0], rate, N)
        assert nerr(one, mo.filt(h, x[0], rate, N, polyorder)) <= tol_for(one.dtype)


def test_errors_and_edge_cases(rng):
    h = rng.random(33).astype(np.float32)
    x = rand_samples(rng, 64, np.float32)
    f = mr.FIRFilter(h, Fraction(3, 4))
    with pytest.raises(mr.MrbError, match="buffer is too small") as e:        # src/Filters.jl:550
        f.filt_(np.empty(3, np.float32), x)
    assert e.value.code == F.MRB_ERR_BUFFER_TOO_SMALL
    with pytest.raises(mr.MrbError, match="buffer length must be >= x length"):   # :460
        mr.FIRFilter(h).filt_(np.empty(3, np.float32), x)
    with pytest.raises(mr.MrbError, match="must be >= interpolation"):        # :503
        mr.FIRFilter(h, Fraction(3, 1)).filt_(np.empty(3, np.float32), x)
    # filt! return conventions (:472,516 buffer ; :574,630,741,835 count)
    buf = np.empty(200, np.float32)
    assert mr.filt_(buf, mr.FIRFilter(h), x) is buf
    assert mr.filt_(buf, mr.FIRFilter(h, Fraction(1, 4)), x) == 16
    assert mr.filt_(buf, f, x) == 48
    # empty input, empty output
    d = mr.FIRFilter(h, Fraction(1, 8))
    assert len(d.filt(x[:0])) == 0 and d.kernel.inputDeficit == 1
    assert len(d.filt(x[:1])) == 1 and d.kernel.inputDeficit == 8
    for i in range(1, 8):
        assert len(d.filt(x[i:i + 1])) == 0 and d.kernel.inputDeficit == 8 - i
    assert len(d.filt(x[8:9])) == 1
    # history carry when the chunk is shorter than the history (shiftin!, src/support.jl:69-76)
    o = mo.FIRFilter(h, Fraction(1, 8))
    for a, b in [(0, 1), (1, 9)]:
        o.filt(x[a:b])
    assert np.array_equal(d.history, o.history[0])
    # one tap (historyLen == 0)
    assert nerr(mr.filt(np.array([2.0]), np.arange(8.0)), 2.0 * np.arange(8.0)) == 0.0


def test_multichannel_matches_per_channel(rng):
    """Channels are independent and share one state machine: a (C, n) batch == C single-channel filters."""
    h = mo.firdes(24 * 147, 0.5 / 147, 7.8562).astype(np.float32)
    x = rand_samples(rng, (37, 3000), np.complex64)
    f = mr.FIRFilter(h, Fraction(147, 160))
    y = np.concatenate([f.filt(x[:, :1111]), f.filt(x[:, 1111:])], axis=1)
    for c in (0, 1, 17, 36):
        assert nerr(y[c], mo.filt(h, x[c], Fraction(147, 160))) <= 1e-5
    assert y.shape == (37, mo.FIRFilter(h, Fraction(147, 160)).outputlength(3000))


@pytest.mark.parametrize("case", ["rational", "decimator", "interpolator", "standard"])
def test_device_path_torch_streaming(case, rng):
    """Zero-copy device path (mrb_filt on torch's stream): state stays on the device between 64K-sample
    chunks; compared with the host path and the oracle; generic and tiled kernels must agree."""
    import torch
    cfg = {"rational": (Fraction(147, 160), 24 * 147, np.complex64, 0.5 / 147),
           "decimator": (Fraction(1, 8), 256, np.complex64, 0.5 / 8),
           "interpolator": (Fraction(4, 1), 128, np.float32, 0.5 / 4),
           "standard": (Fraction(1, 1), 128, np.float32, 0.25)}[case]
    ratio, ntaps, tx, cutoff = cfg
    h = mo.firdes(ntaps, cutoff, 7.8562).astype(np.float32)
    nch, chunk = 70, 1 << 14
    x = rand_samples(rng, (nch, 3 * chunk + 5), tx)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    g.set_kernel_policy(1)                                            # force the generic kernel
    o = mo.FIRFilter(h, ratio)
    edges = [0, chunk, chunk + 1, 2 * chunk + 3, 3 * chunk + 5]
    for a, b in zip(edges[:-1], edges[1:]):
        yd = f.filt(xd[:, a:b])
        yg = g.filt(xd[:, a:b])
        w = o.filt(x[:5, a:b])
        torch.cuda.synchronize()
        assert yd.is_cuda and tuple(yd.shape) == (nch, w.shape[1])
        y = yd.cpu().numpy()
        assert nerr(y[:5], w) <= 1e-5, (case, nerr(y[:5], w))
        assert nerr(yg.cpu().numpy(), y) <= 2e-6
        assert states_equal(f, o)
    assert f.last_kernel != "none" and g.last_kernel == "generic"
    assert f.launch_count > 0


def test_full_size_properties_c5_shard(rng):
    """BASELINE config 5 at full per-GPU chunk size, properties that need no oracle run over the whole
    batch: (i) exact count; (ii) chunking invariance (64K one-shot == 4 x 16K streamed; the first ~24 outputs
    of a chunk come from the generic kernel, whose summation order differs, hence a 1e-6 bound instead of
    bit equality); (iii) linearity in x (exact: scaling by 2); (iv) channel independence (a channel's output
    does not depend on its neighbours); spot rows checked against the oracle."""
    import torch
    h = mo.firdes(24 * 147, 0.5 / 147, 7.8562).astype(np.float32)
    nch, n = 1024, 1 << 16
    gen = torch.Generator(device="cuda"); gen.manual_seed(0x4D520005)
    x = torch.view_as_complex(torch.rand((nch, n, 2), generator=gen, device="cuda"))
    ratio = Fraction(147, 160)
    y1 = mr.FIRFilter(h, ratio).filt(x)
    assert y1.shape[1] == 60212 == mo.FIRFilter(h, ratio).outputlength(n)
    f = mr.FIRFilter(h, ratio)
    y4 = torch.cat([f.filt(x[:, i * 16384:(i + 1) * 16384]) for i in range(4)], dim=1)
    assert (y1 - y4).abs().max().item() <= 1e-6 * y1.abs().max().item()
    y2 = mr.FIRFilter(h, ratio).filt(2 * x)
    assert torch.equal(torch.view_as_real(y2), torch.view_as_real(2 * y1))
    xs = x[100:164].clone()
    ys = mr.FIRFilter(h, ratio).filt(xs)
    assert (ys - y1[100:164]).abs().max().item() <= 1e-5 * y1.abs().max().item()
    rows = [0, 511, 1023]
    w = mo.filt(h, x[rows].cpu().numpy(), ratio)
    assert nerr(y1[rows].cpu().numpy(), w) <= 1e-5


def test_segment_split_matches_stream(rng):
    """Long-stream split: S independent segments, each seeked to its closed-form start state with a
    tap-length halo, reproduce the single-stream output (same counts, values to 1e-6); no collective involved."""
    import torch
    h = mo.firdes(24 * 147, 0.5 / 147, 7.8562).astype(np.float32)
    ratio = Fraction(147, 160)
    n = 200_003
    x = torch.from_numpy(rand_samples(rng, (1, n), np.complex64)).cuda()
    whole = mr.FIRFilter(h, ratio).filt(x)
    bounds = [0, 50_001, 99_999, 160_000, n]
    parts = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        f = mr.FIRFilter(h, ratio, nchannels=1, sample_dtype=np.complex64)
        H = f.historyLen
        k0 = C.c_int64()
        halo = None
        if a > 0:
            halo = x[:, a - H:a].contiguous()
        F.check(F.lib().mrb_seek(f._handle, a, halo.data_ptr() if halo is not None else None, H, C.byref(k0),
                                 torch.cuda.current_stream().cuda_stream))
        assert k0.value == sum(p.shape[1] for p in parts)
        parts.append(f.filt(x[:, a:b]))
    got = torch.cat(parts, dim=1)
    assert got.shape == whole.shape
    assert (got - whole).abs().max().item() <= 1e-6 * whole.abs().max().item()


@pytest.mark.parametrize("case", ["rational_tiled", "standard_unit", "decimator", "arbitrary", "farrow", "rational_f64"])
def test_live_tap_update(case, rng):
    """SURVEY 8f rank 3 (host side): taps replaced between chunks with state and history carried -- every fast kernel's
    private copy of the bank must follow.  Oracle twin: a fresh oracle filter with the new taps that inherits the old
    one's kernel state and history."""
    import torch
    tx, nch, n, cut = np.complex64, 130, 24000, 12000                # cut keeps the second chunk 16-byte aligned
    if case == "rational_tiled":
        mk = lambda: (rng.standard_normal(24 * 7).astype(np.float32), Fraction(7, 8))
    elif case == "standard_unit":
        tx = np.float32
        mk = lambda: (rng.standard_normal(100).astype(np.float32), Fraction(1, 1))
    elif case == "decimator":
        mk = lambda: (rng.standard_normal(200).astype(np.float32), Fraction(1, 8))
    elif case == "arbitrary":
        tx = np.float32
        mk = lambda: (rng.standard_normal(32 * 9).astype(np.float32), 0.918734, 32)
    elif case == "farrow":
        tx = np.float32
        mk = lambda: (rng.standard_normal(32 * 9).astype(np.float32), 0.918734, 32, 3)
    else:
        tx, nch = np.float64, 2
        mk = lambda: (rng.standard_normal(24 * 7), Fraction(7, 8))
    a1, a2 = mk(), mk()
    x = rand_samples(rng, (nch, n), tx)
    xd = torch.from_numpy(x).cuda()
    f, o = mr.FIRFilter(*a1), mo.FIRFilter(*a1)
    y1, w1 = f.filt(xd[:, :cut]), o.filt(x[:3, :cut])
    assert nerr(y1[:3].cpu().numpy(), w1) <= tol_for(y1.cpu().numpy().dtype)
    first_kernel = f.last_kernel
    f.set_taps(a2[0])
    o2 = mo.FIRFilter(*a2)
    o2.history = o.history
    for name in ("phiIdx", "inputDeficit", "phiAccumulator", "alpha", "xIdx"):
        if hasattr(o.kernel, name):
            setattr(o2.kernel, name, getattr(o.kernel, name))
    y2, w2 = f.filt(xd[:, cut:]), o2.filt(x[:3, cut:])
    assert y2.shape[1] == w2.shape[1]
    assert nerr(y2[:3].cpu().numpy(), w2) <= tol_for(y2.cpu().numpy().dtype)
    assert states_equal(f, o2)
    assert f.last_kernel == first_kernel and first_kernel != "generic"


@pytest.mark.parametrize("case", ["farrow", "arbitrary", "rational"])
def test_output_time_offset_api(case, rng):
    """the documented use of setphase (examples/FIRFarrow.jl:25-30) -- throw away whole samples by
    raising kernel.inputDeficit, set the fractional phase with setphase, then filt -- against the oracle with the same
    two pokes: count, state and values; then a second chunk continues from the carried state."""
    N, tpp = 32, 10
    h = (mo.firdes(tpp * N, 0.45 / N, 5.6533) * N)
    if case == "farrow":
        args = (h, 1.1234, N, 4)
    elif case == "arbitrary":
        args = (h, 1.1234, N)
    else:
        args = (h[:96], Fraction(7, 5))
    x = rand_samples(rng, (2, 5000), np.float64)
    delay = (len(args[0]) - 1) / (2 * N) + 3.5                       # examples/FIRFarrow.jl:25-27
    phase, throwaway = np.modf(delay)
    f, o = mr.FIRFilter(*args), mo.FIRFilter(*args)
    f.kernel.inputDeficit += int(throwaway)
    o.kernel.inputDeficit += int(throwaway)
    assert f.setphase(phase) == o.setphase(phase)
    for a, b in ((0, 3), (3, 2600), (2600, 5000)):                   # the first chunk is shorter than the deficit
        y, w = f.filt(x[:, a:b]), o.filt(x[:, a:b])
        assert y.shape == w.shape
        assert nerr(y, w) <= 1e-12
        assert states_equal(f, o)


@pytest.mark.parametrize("polyorder", [None, 4])
@pytest.mark.parametrize("tx", [np.float32, np.float64])
def test_segment_split_arbitrary_and_farrow(polyorder, tx, rng):
    """a long stream through an arbitrary-rate / Farrow resampler split into independent segments.
    Each segment's filter is seeked by exact replay and given the preceding samples as halo; counts, first-output
    indices and end states are exact, values match the single-stream run (and the oracle on a slice)."""
    import torch
    N, rate = 32, 0.918734
    h = (mo.firdes(12 * N, 0.45 / N, 5.6533) * N).astype(tx)
    args = (h, rate, N) if polyorder is None else (h, rate, N, polyorder)
    nch, n = 3, 150_001
    x = rand_samples(rng, (nch, n), tx)
    xd = torch.from_numpy(x).cuda()
    whole_f = mr.FIRFilter(*args)
    whole = whole_f.filt(xd)
    plan = mr.segment_plan(whole_f, n, 4, align=1)
    assert [p[0] for p in plan][0] == 0 and plan[-1][1] == n
    parts = []
    for n0, n1, k0, cnt in plan:
        f = mr.FIRFilter(*args, nchannels=nch, sample_dtype=tx)
        H = f.historyLen
        halo = xd[:, n0 - H:n0].contiguous() if n0 > 0 else None
        assert f.seek(n0, halo) == k0 == sum(p.shape[1] for p in parts)
        y = f.filt(xd[:, n0:n1])
        assert y.shape[1] == cnt
        parts.append(y)
    got = torch.cat(parts, dim=1)
    assert got.shape == whole.shape
    assert (got - whole).abs().max().item() <= tol_for(tx) * whole.abs().max().item()
    s_seg, s_whole = f._get_state(), whole_f._get_state()
    assert (s_seg.phi_idx, s_seg.input_deficit, s_seg.phi_accumulator, s_seg.alpha) == \
           (s_whole.phi_idx, s_whole.input_deficit, s_whole.phi_accumulator, s_whole.alpha)
    o = mo.FIRFilter(*args)
    w = o.filt(x[0, :20000])
    assert nerr(got[0, :len(w)].cpu().numpy(), w) <= tol_for(tx)


@pytest.mark.parametrize("L,ntaps,nch", [(1, 128, 33), (1, 97, 5), (1, 31, 64), (2, 200, 40), (2, 66, 1), (4, 512, 31),
                                          (4, 390, 96)])
@pytest.mark.parametrize("tx", [np.float32, np.complex64])
def test_unit_stride_f32_kernel_shapes(L, ntaps, nch, tx, rng):
    """The float32 standard / interpolator fast path (mrb_unit.cuh): ragged tap counts (zero-padded tap blocks),
    ragged channel counts (TMA clips rows), chunk lengths that are not whole steps, state carried across
    chunks; against the oracle and the generic kernel."""
    import torch
    h = rng.standard_normal(ntaps).astype(np.float32)
    ratio = Fraction(L, 1)
    n = 5000                                                          # row pitch: a multiple of 16 bytes (TMA)
    x = rand_samples(rng, (nch, n), tx)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    f.set_kernel_policy(2)                                            # the CUDA-core fast paths (float32 would take the tensor-core kernel)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, ratio)
    for a, b in ((0, 2048), (2048, 2052), (2052, 4001), (4001, 4008), (4008, n)):   # 4001: misaligned -> generic
        yd = f.filt(xd[:, a:b])
        yg = g.filt(xd[:, a:b])
        w = o.filt(x[: min(nch, 3), a:b])
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert y.shape == (nch, (b - a) * L)
        assert nerr(y[: min(nch, 3)], w) <= 1e-5
        assert nerr(yg.cpu().numpy(), y) <= 2e-6
    # complex64 with <= 24 taps and L = 1 is the tiled kernel's; everything else here is the unit kernel's
    want = "tiled" if (tx == np.complex64 and L == 1 and ntaps <= 24) else ("unit_c64" if tx == np.complex64 else "unit_f32")
    assert f.last_kernel.startswith(want), f.last_kernel


@pytest.mark.parametrize("M,ntaps,nch", [(8, 256, 33), (8, 100, 5), (8, 7, 64), (4, 128, 40), (4, 33, 1), (2, 64, 31),
                                          (2, 19, 96)])
@pytest.mark.parametrize("tx", [np.complex64, np.float32])
def test_decimator_c64_kernel_shapes(M, ntaps, nch, tx, rng):
    """The complex64 decimator fast path (mrb_decim.cuh): ragged tap counts, ragged channel counts, chunk
    lengths that leave every possible input deficit behind (so the window start takes every alignment), empty
    outputs; count / state exact, values against the oracle and the generic kernel."""
    import torch
    h = rng.standard_normal(ntaps).astype(np.float32)
    ratio = Fraction(1, M)
    n = 9000
    x = rand_samples(rng, (nch, n), tx)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, ratio)
    # float32 chunks must start on a multiple of 4 samples for the fast path (16-byte aligned pointers)
    edges = [0, 3000, 3001, 3003, 6000 + M // 2, 6000 + M // 2 + 2, n] if tx == np.complex64 else [0, 3000, 3004, 3012, 6000 + 4, 6000 + 12, n]
    used = set()
    for a, b in zip(edges[:-1], edges[1:]):
        yd = f.filt(xd[:, a:b])
        yg = g.filt(xd[:, a:b])
        w = o.filt(x[: min(nch, 3), a:b])
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert y.shape == (nch, w.shape[1])
        assert nerr(y[: min(nch, 3)], w) <= 1e-5
        assert nerr(yg.cpu().numpy(), y) <= 2e-6
        assert states_equal(f, o)
        used.add(f.last_kernel)
    if tx == np.complex64 or M >= 4:                                # float32 needs M >= 4 (16-byte TMA rows)
        assert any(k.startswith("decim_c64" if tx == np.complex64 else "decim_f32") for k in used), used


@pytest.mark.parametrize("polyorder", [None, 4])
def test_schedule_cache_is_keyed_by_state_and_length(polyorder, rng):
    """Table kinds: the replayed schedule of a count query is reused by the filt call that follows only when (state, n_in)
    are the same -- repeated lengths, another length, a state edit and a reset in between; counts, phase state and values
    must be those of the oracle every time."""
    import torch
    N = 32
    hLen, beta = mo.kaiserlength(0.05, samplerate=N)
    hLen = -(-hLen // N) * N
    h = (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(np.float32)
    nch = 64
    lens = [20000, 20000, 20000, 15000, 20000, 20001, 20001, 9000, 20000, 20000]
    x = rand_samples(rng, (nch, sum(lens)), np.float32)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, 0.918734, N, polyorder)
    o = mo.FIRFilter(h, 0.918734, N, polyorder)
    a = 0
    for i, n in enumerate(lens):
        if i == 5:                                                 # a state edit between two calls (examples/FIRFarrow.jl:29)
            f.kernel.inputDeficit += 3
            o.kernel.inputDeficit += 3
        if i == 8:
            f.reset(); o.reset()
        assert f.outputlength(n) >= 0                              # (the count query replays or adopts first, filt reuses it)
        y = f.filt(xd[:, a:a + n])
        w = o.filt(x[:2, a:a + n])
        torch.cuda.synchronize()
        assert y.shape == (nch, w.shape[1]), (i, y.shape, w.shape)
        assert nerr(y[:2].cpu().numpy(), w) <= 1e-5, (i, nerr(y[:2].cpu().numpy(), w))
        assert states_equal(f, o), i
        a += n


def test_repeated_one_shot_calls_reuse_a_reset_handle(rng):
    """filt(h, x, ratio) keeps the handle of a one-shot call and resets it when the same taps / ratio / layout come again:
    the repeated call must equal a fresh filter's output bit for bit (mrb_reset is a full re-initialisation), for host and
    device inputs, and a different input length or ratio must not be served by a stale state."""
    import torch
    h = mo.firdes(24 * 147, 0.5 / 147, 7.8562).astype(np.float32)
    x = rand_samples(rng, (5000,), np.float32)
    mr.clear_oneshot_cache()
    y1 = mr.filt(h, x, Fraction(147, 160))
    y2 = mr.filt(h, x, Fraction(147, 160))                          # served by the kept handle
    y3 = mr.FIRFilter(h, Fraction(147, 160)).filt(x)
    assert np.array_equal(y1, y2) and np.array_equal(y1, y3)
    assert nerr(y1, mo.filt(h, x, Fraction(147, 160))) <= 1e-5
    y4 = mr.filt(h, x[:3001], Fraction(147, 160))                   # same key, other length: state must start over
    assert np.array_equal(y4, mr.FIRFilter(h, Fraction(147, 160)).filt(x[:3001]))
    xd = torch.from_numpy(x).cuda()
    y5, y6 = mr.filt(h, xd, Fraction(3, 2)), mr.filt(h, xd, Fraction(3, 2))
    assert torch.equal(y5, y6) and nerr(y5.cpu().numpy(), mo.filt(h, x, Fraction(3, 2))) <= 1e-5
    ya = mr.filt(h[:2336] * 32, x, 0.918734, 32, 4)
    yb = mr.filt(h[:2336] * 32, x, 0.918734, 32, 4)
    assert np.array_equal(ya, yb)
    mr.clear_oneshot_cache()


@pytest.mark.parametrize("case", ["decim8", "f64_arbitrary_dmma", "f64_farrow_dmma"])
def test_mbarrier_pipelines_are_deterministic_over_many_launches(case, rng):
    """k_decim8 and the FP64 tensor-core table kernel hand their ring / staging / tap-row buffers around with mbarriers only
    (no CTA barrier; racecheck cannot follow a bulk-copy refill ordered by an mbarrier).  The same stream of chunks twice
    through fresh filters: every chunk must come out BIT-identical -- a race shows up as a differing chunk (tools/
    soak_new_kernels.py is the long form, 2 x 1500 chunks, profiles/r2_soak_new_kernels.txt)."""
    import torch
    N = 32
    if case == "decim8":
        ratio, h, extra, dt, tdt, want = Fraction(1, 8), mo.firdes(256, 0.5 / 8, 7.8562).astype(np.float32), (), np.complex64, torch.complex64, "decim8_c64"
    else:
        hLen, beta = mo.kaiserlength(0.05, samplerate=N)
        hLen = -(-hLen // N) * N
        h = (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(np.float64)
        ratio, extra, dt, tdt, want = 0.918734, ((N,) if case == "f64_arbitrary_dmma" else (N, 4)), np.float64, torch.float64, "table_f64_dmma"
    nch, n, steps = 512, 32768, 150
    torch.manual_seed(11)
    xs = [torch.randn((nch, n), device="cuda", dtype=tdt) for _ in range(3)]
    sums = []
    for _ in range(2):
        f = mr.FIRFilter(h, ratio, *extra, nchannels=nch, sample_dtype=dt)
        cs = []
        for i in range(steps):
            y = f.filt(xs[i % 3])
            bits = torch.view_as_real(y).view(torch.int32) if tdt == torch.complex64 else y.view(torch.int64)
            cs.append(bits.to(torch.int64).sum())
        torch.cuda.synchronize()
        assert f.last_kernel == want, f.last_kernel
        sums.append(torch.stack(cs).cpu())
    assert int((sums[0] != sums[1]).sum()) == 0


@pytest.mark.parametrize("ntaps,nch", [(256, 300), (255, 129), (100, 512), (9, 160)])
def test_decimator_m8_lane_per_channel_kernel(ntaps, nch, rng):
    """k_decim8 (mrb_decim.cuh): 1//8 on complex64 with the taps as launch constants -- ragged tap and channel counts, chunk
    lengths that leave every input deficit behind (both window alignments), tiles that end inside a step, state carry;
    against the oracle and the generic kernel."""
    import torch
    h = rng.standard_normal(ntaps).astype(np.float32)
    ratio = Fraction(1, 8)
    n = 70000
    x = rand_samples(rng, (nch, n), np.complex64)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=np.complex64)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, ratio)
    edges = [0, 30000, 30001, 30003, 50004, 50010, n]
    rows = [0, 1, nch // 2, nch - 1]
    used = set()
    for a, b in zip(edges[:-1], edges[1:]):
        yd = f.filt(xd[:, a:b])
        yg = g.filt(xd[:, a:b])
        w = o.filt(x[rows, a:b])
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert y.shape == (nch, w.shape[1])
        assert nerr(y[rows], w) <= 1e-5
        assert nerr(yg.cpu().numpy(), y) <= 2e-6
        assert states_equal(f, o)
        used.add(f.last_kernel)
    assert "decim8_c64" in used, used


@pytest.mark.parametrize("th,tx", [(np.float32, np.float32), (np.float32, np.complex64), (np.float64, np.float64),
                                   (np.float64, np.complex128), (np.float64, np.complex64)])
@pytest.mark.parametrize("M,ntaps,nch", [(8, 256, 40), (4, 77, 3), (16, 300, 1)])
def test_head_kernel_short_chunks(th, tx, M, ntaps, nch, rng):
    """k_head_warp (mrb_kernels.cuh): short launches of decimating filters with long windows -- chunk heads, and whole
    chunks when they are short -- one warp per output and channel.  Every sample dtype pairing, chunks from one sample to
    a few hundred outputs, state carried; against the oracle and k_generic."""
    import torch
    h = rng.standard_normal(ntaps).astype(th)
    ratio = Fraction(1, M)
    edges = [0, 1, 2, M + 3, 40 * M + 1, 41 * M, 300 * M + 5, 300 * M + 6, 420 * M]
    x = rand_samples(rng, (nch, edges[-1]), tx)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, ratio)
    used = set()
    for a, b in zip(edges[:-1], edges[1:]):
        y, yg, w = f.filt(xd[:, a:b]).cpu().numpy(), g.filt(xd[:, a:b]).cpu().numpy(), o.filt(x[:, a:b])
        assert y.shape == w.shape
        assert nerr(y, w) <= tol_for(y.dtype) and nerr(yg, y) <= tol_for(y.dtype)
        assert states_equal(f, o)
        if y.shape[1]:
            used.add(f.last_kernel)
    assert "head" in used, used


@pytest.mark.parametrize("ratio,ntaps", [(Fraction(147, 160), 3528), (Fraction(3, 17), 120), (Fraction(1, 1), 63),
                                         (Fraction(7, 1), 100), (Fraction(1, 5), 41), (Fraction(160, 147), 1600)])
@pytest.mark.parametrize("th,tx,nch", [(np.float32, np.float32, 1), (np.float32, np.complex64, 3), (np.float64, np.float64, 2),
                                       (np.float64, np.complex128, 1), (np.float64, np.float32, 1)])
def test_stream_kernel_few_channels(ratio, ntaps, th, tx, nch, rng):
    """k_stream (mrb_kernels.cuh): integer schedules no tiled kernel covers (few channels, Float64, README benchmark
    shape), polyphase bank staged in shared memory.  Chunks long enough to reach it and short ones that stay on
    k_generic, odd chunk boundaries, state carried; count / state exact, values against the oracle and k_generic."""
    import torch
    h = rng.standard_normal(ntaps).astype(th)
    big = int(9000 / ratio) + 1                                     # inputs for > 8192 outputs (k_stream's threshold)
    edges = [0, big, big + 2, big + 499, 2 * big + 499]
    n = edges[-1]
    x = rand_samples(rng, (nch, n), tx)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, ratio)
    used = set()
    for a, b in zip(edges[:-1], edges[1:]):
        yd = f.filt(xd[:, a:b])
        yg = g.filt(xd[:, a:b])
        w = o.filt(x[:, a:b])
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert y.shape == w.shape
        assert nerr(y, w) <= tol_for(y.dtype)
        assert nerr(yg.cpu().numpy(), y) <= tol_for(y.dtype)
        assert states_equal(f, o)
        if y.shape[1] >= 8192:
            used.add(f.last_kernel)
        assert g.last_kernel in ("generic", "none")
    # the long chunks never fall back to k_generic; with a few channels Float64 work runs on k_stream (from 32 channels on:
    # the FP64 tensor-core kernel, test_float64_integer_kinds_on_fp64_tensor_cores)
    assert used and "generic" not in used, used
    if th == np.float64:
        assert used == {"stream"}, used


@pytest.mark.parametrize("ratio,ntaps", [(Fraction(147, 160), 3528), (Fraction(4, 1), 128), (Fraction(1, 1), 63), (Fraction(3, 2), 90),
                                         (Fraction(2, 3), 50), (Fraction(160, 147), 3200)])
@pytest.mark.parametrize("th", [np.float64, np.float32])
def test_float64_integer_kinds_on_fp64_tensor_cores(ratio, ntaps, th, rng):
    """Standard / interpolator / rational filters on Float64 samples: the table kernel's mma.sync.m8n8k4 variant with the
    schedule in closed form (mrb_table.cuh k_table_rows, sn == nullptr).  Ragged channel count, chunk edges that leave every
    phase / deficit behind, state carried; against the oracle and k_generic."""
    import torch
    h = rng.standard_normal(ntaps).astype(th)
    nch = 70
    big = (int(9000 / ratio) + 2) // 2 * 2                          # even: 16-byte aligned chunk starts and row pitch
    # (the chunk that starts at the odd edge takes the unaligned fallback, k_stream)
    edges = [0, big, big + 2, big + 502, 2 * big + 502, 2 * big + 503, 3 * big + 700]
    n = edges[-1]
    x = rand_samples(rng, (nch, n), np.float64)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=np.float64)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=np.float64)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, ratio)
    rows = [0, 1, 35, nch - 1]
    used = set()
    for a, b in zip(edges[:-1], edges[1:]):
        yd = f.filt(xd[:, a:b])
        yg = g.filt(xd[:, a:b])
        w = o.filt(x[rows, a:b])
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert y.shape == (nch, w.shape[1])
        assert nerr(y[rows], w) <= tol_for(y.dtype), (a, b, nerr(y[rows], w))
        assert nerr(yg.cpu().numpy(), y) <= 1e-13
        assert states_equal(f, o)
        used.add(f.last_kernel)
    assert "int_f64_dmma" in used, used
    # host buffers of odd length: mrb_filt_host stages them with 16-byte row pitches, so the fast path still applies
    fh = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=np.float64)
    oh = mo.FIRFilter(h, ratio)
    m = big + 501
    yh = fh.filt(np.ascontiguousarray(x[:, :m]))
    assert nerr(np.asarray(yh)[rows], oh.filt(x[rows, :m])) <= tol_for(np.float64)
    assert fh.last_kernel == "int_f64_dmma", fh.last_kernel


@pytest.mark.parametrize("tx", [np.float32, np.float64, np.complex64])
@pytest.mark.parametrize("polyorder", [None, 4])
@pytest.mark.parametrize("rate,nch", [(0.918734, 33), (1.37, 64), (1 / 2.123456789, 5)])
def test_table_kernel_arbitrary_farrow(tx, polyorder, rate, nch, rng):
    """The arbitrary / farrow fast path (mrb_table.cuh: tap rows built once per chunk, TMA ring, one dot product
    per output): chunked streaming on the device against the oracle and the generic kernel; counts, phase
    accumulator, alpha and deficit bit-exact after every chunk."""
    import torch
    N = 32
    hLen, beta = mo.kaiserlength(0.05, samplerate=N)
    hLen = -(-hLen // N) * N
    th = np.float64 if tx == np.float64 else np.float32
    h = (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(th)          # test/runtests.jl:336-341
    n = 12000
    x = rand_samples(rng, (nch, n), tx)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, rate, N, polyorder)
    g = mr.FIRFilter(h, rate, N, polyorder, nchannels=nch, sample_dtype=tx)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, rate, N, polyorder)
    used = set()
    for a, b in ((0, 5000), (5000, 5004), (5004, 9000), (9000, n)):
        yd = f.filt(xd[:, a:b])
        yg = g.filt(xd[:, a:b])
        w = o.filt(x[:2, a:b])
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert y.shape == (nch, w.shape[1])
        assert nerr(y[:2], w) <= tol_for(tx), (a, b, nerr(y[:2], w))
        # float32: the tensor-core path (3xTF32, tensor-core accumulation order) measures <= 1.6e-6 against k_generic
        assert nerr(yg.cpu().numpy(), y) <= (1e-13 if tx == np.float64 else 4e-6)
        assert states_equal(f, o)
        used.add(f.last_kernel)
    # (float64 below rate 0.5: a step's windows do not fit the shared-memory ring -> generic kernel, by design)
    # float32: the tensor-core kernel (mrb_mma.cuh); float64 / complex64: the table kernel
    assert any(k.startswith(("table_", "mma_")) for k in used) or (tx != np.float32 and rate < 0.5), used
    if tx == np.float64 and rate >= 0.5:
        assert "table_f64_dmma" in used, used                      # the FP64 tensor-core variant (mma.sync m8n8k4)


@pytest.mark.parametrize("case", ["rational", "decimator", "interpolator", "standard", "arbitrary", "farrow"])
def test_host_path_uses_fast_kernels_for_any_length(case, rng):
    """mrb_filt_host stages host buffers with 16-byte row pitches, so odd chunk lengths (and odd output counts)
    still run the TMA fast paths; values against the oracle."""
    N = 32
    hl, beta = mo.kaiserlength(0.05, samplerate=N)
    ha = (mo.firdes(-(-hl // N) * N, 0.45, beta, samplerate=32) * N).astype(np.float32)
    cfg = {"rational": (Fraction(147, 160), mo.firdes(24 * 147, 0.5 / 147, 7.8562).astype(np.float32), np.complex64, "mma_c64"),
           "decimator": (Fraction(1, 8), mo.firdes(256, 0.5 / 8, 7.8562).astype(np.float32), np.complex64, "decim"),
           "interpolator": (Fraction(4, 1), mo.firdes(128, 0.5 / 4, 7.8562).astype(np.float32), np.float32, "mma"),
           "standard": (Fraction(1, 1), mo.firdes(128, 0.25, 7.8562).astype(np.float32), np.float32, "mma"),
           "arbitrary": (0.918734, ha, np.float32, "mma"), "farrow": (0.918734, ha, np.float32, "mma")}[case]
    ratio, h, tx, want = cfg
    extra = (N, 4) if case == "farrow" else (N,) if case == "arbitrary" else ()
    x = rand_samples(rng, (64, 7001), tx)
    f, o = mr.FIRFilter(h, ratio, *extra), mo.FIRFilter(h, ratio, *extra)
    for a, b in ((0, 3001), (3001, 7001)):
        y, w = f.filt(x[:, a:b]), o.filt(x[:3, a:b])
        assert y.shape == (64, w.shape[1])
        assert nerr(y[:3], w) <= 1e-5
        assert f.last_kernel.startswith(want), (case, f.last_kernel)
        assert states_equal(f, o)


def test_fast_paths_random_stress_against_generic_kernel(rng):
    """Randomised shapes for every fast path, compared with the generic kernel (itself pinned to the oracle above)
    on the same device data: ratios, tap counts, channel counts, chunkings.  Shakes out ring / schedule corner
    cases (tile ends, prefetch bounds, alignment fallbacks); counts and carried state must agree exactly."""
    import math
    import torch
    r = np.random.default_rng(0x4D52 + 99)
    cases = []
    for _ in range(14):                                            # tiled: rational L <= M < 2L, complex64, T <= 24
        L = int(r.integers(2, 200))
        M = int(r.integers(L, 2 * L))
        while math.gcd(L, M) != 1:
            M = int(r.integers(L, 2 * L))
        cases.append((Fraction(L, M), int(r.integers(1, 24 * L + 1)), np.complex64, "tiled"))
    cases.append((Fraction(1, 1), 24, np.complex64, "tiled"))
    cases.append((Fraction(160, 147), 24 * 160, np.complex64, "tiled"))         # interpolating rationals: M < L <= 1.5 M
    for _ in range(8):
        M = int(r.integers(2, 150))
        L = int(r.integers(M + 1, M + M // 2 + 2))
        while math.gcd(L, M) != 1 or 3 * (L - M) > L:
            L = int(r.integers(M + 1, M + M // 2 + 2))
        cases.append((Fraction(L, M), int(r.integers(1, 24 * L + 1)), np.complex64, "tiled"))
    for _ in range(4):                                             # unit: float32, L in {1, 2, 4}
        L = int(r.choice([1, 2, 4]))
        cases.append((Fraction(L, 1), int(r.integers(1, 128 * L + 1)), np.float32, "unit"))
    for _ in range(4):                                             # unit: complex64, more than 24 taps per phase
        L = int(r.choice([1, 2, 4]))
        cases.append((Fraction(L, 1), int(r.integers(24 * L + 1, 128 * L + 1)), np.complex64, "unit_c64"))
    for _ in range(4):                                             # decimator: complex64, M in {2, 4, 8}
        M = int(r.choice([2, 4, 8]))
        cases.append((Fraction(1, M), int(r.integers(1, 32 * M + 1)), np.complex64, "decim"))
    for _ in range(3):                                             # decimator: float32, M in {4, 8}
        M = int(r.choice([4, 8]))
        cases.append((Fraction(1, M), int(r.integers(1, 32 * M + 1)), np.float32, "decim_f32"))
    for _ in range(4):                                             # tensor cores: complex64 rational through the float view
        M = int(r.integers(2, 60))
        L = int(r.integers(M // 2 + 1, 2 * M))
        while math.gcd(L, M) != 1 or L == M:
            L = int(r.integers(M // 2 + 1, 2 * M))
        cases.append((Fraction(L, M), int(r.integers(1, 40 * L + 1)), np.complex64, "mma_c64"))
    for _ in range(4):                                             # tensor cores, split mode: complex64 standard / interpolator
        L = int(r.choice([1, 2, 3, 4, 7]))
        cases.append((Fraction(L, 1), int(r.integers(25 * L, 70 * L + 1)), np.complex64, "mma_c64"))
    for ratio, ntaps, tx, want in cases:
        h = r.standard_normal(ntaps).astype(np.float32)
        nch = int(r.integers(48, 200)) if want == "mma_c64" else int(r.integers(1, 200))
        n = 4 * int(r.integers(1500, 3000))
        x = torch.from_numpy(rand_samples(r, (nch, n), tx)).cuda()
        f = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
        if want in ("unit", "unit_c64", "decim_f32", "tiled"):
            f.set_kernel_policy(2)                                 # keep to the CUDA-core fast paths here (tensor cores: below)
        g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=tx)
        g.set_kernel_policy(1)
        cut = sorted(4 * int(v) for v in r.integers(1, n // 4, size=2))
        used = set()
        for a, b in zip([0] + cut, cut + [n]):
            yf, yg = f.filt(x[:, a:b]), g.filt(x[:, a:b])
            assert yf.shape == yg.shape, (ratio, ntaps, nch, a, b)
            if yf.numel():
                err = (yf - yg).abs().max().item() / max(yg.abs().max().item(), 1e-30)
                assert err <= 5e-6, (ratio, ntaps, nch, a, b, err, f.last_kernel)
            sf, sg = f._get_state(), g._get_state()
            assert (sf.phi_idx, sf.input_deficit) == (sg.phi_idx, sg.input_deficit)
            used.add(f.last_kernel)
        torch.cuda.synchronize()
        assert any(k.startswith(want) for k in used) or n < 2000, (ratio, ntaps, nch, used)


@pytest.mark.parametrize("case", ["rational", "decimator", "interpolator"])
def test_long_stream_in_place_segmentation(case, rng):
    """filt_long_stream: one long stream viewed in place as a matrix of M-aligned segments (halo = the samples
    that precede each segment in memory) equals the plain single-channel call; counts exact."""
    import torch
    cfg = {"rational": (Fraction(147, 160), mo.firdes(24 * 147, 0.5 / 147, 7.8562).astype(np.float32), np.complex64),
           "decimator": (Fraction(1, 8), mo.firdes(256, 0.5 / 8, 7.8562).astype(np.float32), np.complex64),
           "interpolator": (Fraction(4, 1), mo.firdes(128, 0.5 / 4, 7.8562).astype(np.float32), np.float32)}[case]
    ratio, h, tx = cfg
    n = 1_500_017
    x = torch.from_numpy(rand_samples(rng, (n,), tx)).cuda()
    whole = mr.FIRFilter(h, ratio).filt(x)
    got = mr.filt_long_stream(h, ratio, x, rows_target=96)
    assert got.shape == whole.shape
    assert (got - whole).abs().max().item() <= 2e-6 * whole.abs().max().item()
    w = mo.filt(h, x[:30000].cpu().numpy(), ratio)
    assert nerr(got[:w.shape[0]].cpu().numpy(), w) <= 1e-5


def _c4_taps(th):
    N = 32
    hLen, beta = mo.kaiserlength(0.05, samplerate=N)
    hLen = -(-hLen // N) * N
    return (mo.firdes(hLen, 0.45, beta, samplerate=32) * N).astype(th)          # test/runtests.jl:336-341


def _c_oracle(h, tx, nch, rate, polyorder):
    import c_oracle as co
    pn = None
    if polyorder is not None:
        pn = mr.pfb2pnfb(mr.taps2pfb(h, 32), polyorder)      # coefficients as data: the library's fit on both sides
    return co.COracleFilter("farrow" if polyorder is not None else "arbitrary", h, tx, nch, rate=rate, Nphi=32,
                            polyorder=polyorder or 0, pnfb=pn)


@pytest.mark.parametrize("tx", [np.float32, np.float64, np.complex64])
@pytest.mark.parametrize("polyorder", [None, 4])
@pytest.mark.parametrize("n_in", [80_000, 160_000])
def test_host_pipeline_multi_slice_arbitrary_farrow(tx, polyorder, n_in, rng):
    """Regression for the round-1 race (VERDICT weak #1): mrb_filt_host pipelines channel blocks over several streams,
    and with more than 65,536 outputs per call the arbitrary-rate kinds upload their schedule in several slices.  Every
    stream must own the buffers it writes.  Pinned numpy input (truly asynchronous copies), >= 3 channel blocks on 3
    streams, 2 (80,000 inputs) and 3 (160,000 inputs) schedule slices: all channels against the device path, 8 channels
    against the C oracle, carried state bit-exact; twice, so that slot reuse across calls is covered too."""
    import torch
    th = np.float64 if tx == np.float64 else np.float32
    h = _c4_taps(th)
    rate = 0.918734
    nch = 24
    xt = torch.empty((nch, 2 * n_in), dtype=getattr(torch, np.dtype(tx).name)).pin_memory()
    x = xt.numpy()
    x[...] = rand_samples(rng, (nch, 2 * n_in), tx)
    f = mr.FIRFilter(h, rate, 32, polyorder, nchannels=nch, sample_dtype=tx)
    f.set_host_pipeline(1, 3)                                    # 1-MiB blocks -> 1..3 channels per block, 3 streams
    g = mr.FIRFilter(h, rate, 32, polyorder, nchannels=nch, sample_dtype=tx)     # device path, one stream
    o = _c_oracle(h, tx, 8, rate, polyorder)
    xd = torch.from_numpy(x).cuda()
    for a, b in ((0, n_in), (n_in, 2 * n_in)):
        N = f._exact_count(b - a)
        assert N > 65536 * (1 if n_in == 80_000 else 2)
        yt = torch.empty((nch, N), dtype=xt.dtype).pin_memory()
        y = yt.numpy()
        assert f.filt_(y, x[:, a:b]) == N
        yd = g.filt(xd[:, a:b].contiguous())
        torch.cuda.synchronize()
        w = o.filt(x[:8, a:b])
        assert w.shape == (8, N)
        assert nerr(y[:8], w) <= tol_for(tx), nerr(y[:8], w)
        # host path and device path run the same kernels on the same data: identical results
        assert np.array_equal(y, yd.cpu().numpy())
        s, so = f._get_state(), o.state()
        assert (s.input_deficit, s.phi_accumulator) == (so["inputDeficit"], so["acc"])
        if polyorder is None:
            assert (s.phi_idx, s.alpha) == (so["phiIdx"], so["alpha"])
        sg = g._get_state()
        assert (s.phi_idx, s.input_deficit, s.phi_accumulator, s.alpha) == (sg.phi_idx, sg.input_deficit, sg.phi_accumulator, sg.alpha)
    assert f.last_kernel.startswith("table_") or f.last_kernel.startswith("mma_"), f.last_kernel


@pytest.mark.parametrize("tx", [np.float32, np.float64])
@pytest.mark.parametrize("polyorder", [None, 4])
def test_full_size_c4_against_oracle(tx, polyorder, rng):
    """BASELINE configs[3] at full size: 1024 channels x 65,536-sample chunks, arbitrary and farrow, Float32 and
    Float64, two streamed chunks; 8 sampled channels against the C oracle, every channel against channel-independence
    (channels 0 and 512 carry the same samples), state bit-exact."""
    import torch
    th = np.float64 if tx == np.float64 else np.float32
    h = _c4_taps(th)
    nch, n = 1024, 65536
    pick = [0, 1, 255, 511, 512, 700, 1000, 1023]
    x = rand_samples(rng, (nch, 2 * n), tx)
    x[512] = x[0]
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, 0.918734, 32, polyorder, nchannels=nch, sample_dtype=tx)
    o = _c_oracle(h, tx, len(pick), 0.918734, polyorder)
    for a, b in ((0, n), (n, 2 * n)):
        y = f.filt(xd[:, a:b].contiguous())
        torch.cuda.synchronize()
        w = o.filt(x[pick, a:b])
        yh = y.cpu().numpy()
        assert yh.shape == (nch, w.shape[1])
        assert nerr(yh[pick], w) <= tol_for(tx), nerr(yh[pick], w)
        assert np.array_equal(yh[0], yh[512])
        s, so = f._get_state(), o.state()
        assert (s.input_deficit, s.phi_accumulator) == (so["inputDeficit"], so["acc"])


def test_setphase_one_before_first_filt(rng):
    """setphase(1.0) on an unbound arbitrary filter leaves the accumulator at Nphi+1; the rebind at the
    first filt must carry that state over (mrb_set_state accepts the closed bound) and match the oracle."""
    h = _c4_taps(np.float32)
    x = rand_samples(rng, 3000, np.float32)
    f, o = mr.FIRFilter(h, 0.918734, 32), mo.FIRFilter(h, 0.918734, 32)
    f.setphase(1.0)
    o.setphase(1.0)
    y, w = f.filt(x), o.filt(x)
    assert y.shape == w.shape and nerr(y, w) <= 1e-5
    assert states_equal(f, o)


@pytest.mark.parametrize("ratio,ntaps,nch", [(Fraction(1, 1), 128, 140), (Fraction(1, 1), 150, 50), (Fraction(1, 1), 60, 128), (Fraction(4, 1), 128, 63),
                                              (Fraction(5, 1), 300, 200), (Fraction(147, 160), 3528, 129), (Fraction(160, 147), 3840, 64),
                                              (Fraction(3, 2), 100, 49), (Fraction(5, 7), 333, 300), (Fraction(2, 1), 9, 48), (Fraction(1, 2), 100, 77),
                                              (Fraction(1, 3), 40, 256)])
def test_tensor_core_kernel_integer_ratios(ratio, ntaps, nch, rng):
    """mrb_mma.cuh on the integer-ratio kinds (float32 samples and taps): standard, interpolator, rational and gentle
    decimators run as banded tcgen05 products with 3xTF32 (the K5 experiment of SURVEY 7).  Periodic schedules build one
    period of tap tiles (ONE resident tile for standard / interpolators).  Ragged taps and channels, streamed chunks with
    carried phase / deficit and a chunk head that reaches into the history, odd chunk lengths (-> other kernels): against
    the oracle (float64 accumulation) and the generic kernel; counts and state exact."""
    import torch
    h = rng.standard_normal(ntaps).astype(np.float32)
    n = 9000
    x = rand_samples(rng, (nch, n), np.float32)
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=np.float32)
    g = mr.FIRFilter(h, ratio, nchannels=nch, sample_dtype=np.float32)
    g.set_kernel_policy(1)
    o = mo.FIRFilter(h, ratio)
    used = set()
    for a, b in ((0, 4000), (4000, 4004), (4004, 6001), (6001, 6008), (6008, n)):
        yd, yg = f.filt(xd[:, a:b]), g.filt(xd[:, a:b])
        w = o.filt(x[: min(nch, 3), a:b])
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        assert y.shape == (nch, w.shape[1])
        if w.size:
            assert nerr(y[: min(nch, 3)], w) <= 1e-5, (a, b, nerr(y[: min(nch, 3)], w), f.last_kernel)
            assert nerr(yg.cpu().numpy(), y) <= 4e-6, (a, b, f.last_kernel)
        assert states_equal(f, o)
        used.add(f.last_kernel)
    assert any(k.startswith("mma_") for k in used), used


def test_live_tap_update_is_asynchronous_and_cheap(rng):
    """SURVEY 8f rank 3 / VERDICT r1 #9: a tap update between 64K-sample chunks is stream ordered (no device-wide
    synchronisation), rebuilds the global-memory banks on the device, and costs tens of microseconds of host time.
    Chunk i is filtered with taps i even though chunk i-1 is still in flight when the taps change; checked against
    fresh filters seeded with the carried state, for a parameter-block kernel (tiled), a global-bank kernel (tensor-core /
    generic) and the decimator tables."""
    import time
    import torch
    cases = [(Fraction(147, 160), 24 * 147, np.complex64), (Fraction(147, 160), 24 * 147, np.float32), (Fraction(1, 8), 256, np.complex64)]
    for ratio, ntaps, tx in cases:
        nch, n = 256, 65536
        hs = [rng.standard_normal(ntaps).astype(np.float32) for _ in range(4)]
        x = torch.from_numpy(rand_samples(rng, (nch, n), tx)).cuda()
        f = mr.FIRFilter(hs[0], ratio, nchannels=nch, sample_dtype=tx)
        ys, cost = [], []
        for i in range(4):
            if i:
                t0 = time.perf_counter()
                f.set_taps(hs[i])                                    # chunk i-1 is still running on the stream
                cost.append(time.perf_counter() - t0)
            ys.append(f.filt(x))
        torch.cuda.synchronize()
        # reference: the same chunks through filters that carry the state explicitly and are BUILT with taps i
        g = mr.FIRFilter(hs[0], ratio, nchannels=nch, sample_dtype=tx)
        for i in range(4):
            if i:
                st, hist = g._get_state(), g.history
                g = mr.FIRFilter(hs[i], ratio, nchannels=nch, sample_dtype=tx)
                g._set_state(st)
                g.history = hist
            w = g.filt(x)
            torch.cuda.synchronize()
            assert ys[i].shape == w.shape
            err = (ys[i] - w).abs().max().item() / w.abs().max().item()
            assert err <= 4e-6, (ratio, tx, i, err, f.last_kernel)
        assert np.median(cost) < 200e-6, (ratio, tx, cost)          # measured ~15-30 us; the bound is loose for noisy hosts
        print("set_taps host cost (us):", [round(c * 1e6, 1) for c in cost], f.last_kernel)


@pytest.mark.parametrize("case", ["rational_c64_tiled", "standard_f32_mma", "interp_f32_mma", "decim_c64", "arbitrary_f32_mma", "arbitrary_f64_table"])
def test_non_finite_sample_poisons_a_bounded_neighbourhood(case, rng):
    """Known deviation: the fast paths pad tap rows with zeros, and 0 * Inf = NaN, so ONE non-finite
    sample can poison outputs whose reference window does not contain it.  This pins the extent: the generic kernel
    (policy 1) is non-finite exactly where the reference's windows contain the sample; a fast path is non-finite on a
    superset of those outputs that exceeds it by no more than the padding (a couple of tap blocks around it), and equals
    the generic kernel everywhere else."""
    import torch
    N = 32
    cfg = {"rational_c64_tiled": (Fraction(147, 160), 24 * 147, np.complex64, (), 64),
           "standard_f32_mma": (Fraction(1, 1), 100, np.float32, (), 256),
           "interp_f32_mma": (Fraction(4, 1), 128, np.float32, (), 4 * 64),
           "decim_c64": (Fraction(1, 8), 256, np.complex64, (), 80),
           "arbitrary_f32_mma": (0.918734, 2336, np.float32, (N,), 256),
           "arbitrary_f64_table": (0.918734, 2336, np.float64, (N,), 256)}[case]
    ratio, ntaps, tx, extra, slack = cfg
    th = np.float64 if tx == np.float64 else np.float32
    h = rng.standard_normal(ntaps).astype(th)
    nch, n, bad = 64, 20000, 9001
    x = rand_samples(rng, (nch, n), tx)
    x[5, bad] = np.inf
    xd = torch.from_numpy(x).cuda()
    f = mr.FIRFilter(h, ratio, *extra, nchannels=nch, sample_dtype=tx)
    g = mr.FIRFilter(h, ratio, *extra, nchannels=nch, sample_dtype=tx)
    g.set_kernel_policy(1)
    y, w = f.filt(xd).cpu().numpy(), g.filt(xd).cpu().numpy()
    assert f.last_kernel != "generic", f.last_kernel
    nf_fast, nf_ref = ~np.isfinite(y), ~np.isfinite(w)
    assert not nf_ref[np.arange(nch) != 5].any() and not nf_fast[np.arange(nch) != 5].any()     # other channels untouched
    kr = np.flatnonzero(nf_ref[5]); kf = np.flatnonzero(nf_fast[5])
    assert kr.size > 0 and set(kr) <= set(kf), (case, kr[:3], kf[:3])
    assert kf.min() >= kr.min() - slack and kf.max() <= kr.max() + slack, (case, kr.min(), kr.max(), kf.min(), kf.max())
    ok = ~nf_fast
    assert nerr(y[ok], w[ok]) <= (1e-12 if tx == np.float64 else 4e-6)
