"""Host-side mirror of Multirate.jl's operator interface for the streaming polyphase FIR path.

Same names, argument meaning and error behaviour as the reference's exports
(src/Multirate.jl:26-41): FIRFilter, FIRStandard/FIRInterpolator/FIRDecimator/
FIRRational/FIRArbitrary/FIRFarrow, filt, filt! (spelled `filt_`), reset, setphase,
outputlength, inputlength, taps2pfb, tapsforphase, tapsforphase! (`tapsforphase_`).
All compute goes through the C-ABI of include/mrb.h (libmrb.so, sm_100a); this file
holds no arithmetic on samples.

One addition: x may be a matrix.  numpy / torch arrays are row-major, so a batch
is shaped (n_channels, n_samples) -- the same memory layout as the Julia Matrix of
size (n_samples, n_channels) -- and every channel shares one state machine.

 - numpy input  -> host path (mrb_filt_host: staged H2D, kernels, D2H), numpy output
 - torch CUDA tensor input -> device path (mrb_filt on torch's current stream, zero copy)
"""
from __future__ import annotations

import ctypes as C
from fractions import Fraction

import numpy as np

from . import _ffi

_DT = {np.dtype(np.float32): _ffi.F32, np.dtype(np.float64): _ffi.F64,
       np.dtype(np.complex64): _ffi.C64, np.dtype(np.complex128): _ffi.C128}
_NP = {v: k for k, v in _DT.items()}


# --------------------------------------------------------------------------------------
# host utilities (construction-time only; src/Filters.jl:284-321, src/support.jl:85-88)
# --------------------------------------------------------------------------------------
def taps2pfb(h, Nphi):
    """taps2pfb(h, Nphi), src/Filters.jl:284-298.  Returns the (tapsPerphi, Nphi) matrix."""
    h = np.ascontiguousarray(h)
    if h.dtype not in (np.float32, np.float64):
        h = h.astype(np.float64)
    T = -(-len(h) // int(Nphi))
    out = np.empty((int(Nphi), T), dtype=h.dtype)        # phase-major == Julia column-major T x Nphi
    _ffi.check(_ffi.lib().mrb_taps2pfb(h.ctypes.data, len(h), _DT[h.dtype], int(Nphi), out.ctypes.data))
    return out.T


def polyfit(y, polyorder):
    """polyfit(y, order), src/support.jl:85-88: least squares on the Vandermonde matrix of x = 1..len(y),
    solved in Float64; coefficients lowest order first."""
    y = np.asarray(y, dtype=np.float64)
    A = np.vander(np.arange(1, len(y) + 1, dtype=np.float64), polyorder + 1, increasing=True)
    return np.linalg.lstsq(A, y, rcond=None)[0]


def pfb2pnfb(pfb, polyorder):
    """pfb2pnfb(pfb, order), src/Filters.jl:311-321: one polynomial per tap row, stored as Poly{T}
    (coefficients rounded to the tap type).  Returned as float64 (tapsPerphi, order+1).
    The fit is the library's (mrb_pfb2pnfb: Householder QR in Float64), the one recipe every binding shares --
    the problem is ill-conditioned enough that numpy's SVD solve and Julia's QR differ in the 10th digit."""
    pfb = np.asarray(pfb)
    if pfb.dtype not in (np.float32, np.float64):
        pfb = pfb.astype(np.float64)
    T, Nphi = pfb.shape
    # the bank back to the tap vector it came from (taps2pfb, src/Filters.jl:284-298): h[r*Nphi + c] = pfb[T-1-r, c]
    h = np.ascontiguousarray(pfb[::-1, :].reshape(-1))
    out = np.empty((T, int(polyorder) + 1), dtype=np.float64)
    _ffi.check(_ffi.lib().mrb_pfb2pnfb(h.ctypes.data, len(h), _DT[h.dtype], int(Nphi), int(polyorder), out.ctypes.data))
    return out


def nextphase(currentphase, ratio):
    """nextphase(currentphase, ratio), src/Filters.jl:433-439 (1-based)."""
    ratio = Fraction(ratio)
    out = C.c_int64()
    _ffi.check(_ffi.lib().mrb_nextphase(int(currentphase), ratio.numerator, ratio.denominator, C.byref(out)))
    return out.value


# --------------------------------------------------------------------------------------
# kernel views: the reference's kernel structs (src/Filters.jl:15-147) as live views of the handle
# --------------------------------------------------------------------------------------
class FIRKernel:
    def __init__(self, owner):
        self._o = owner

    def _st(self):
        return self._o._get_state()

    def _set(self, **kw):
        s = self._o._get_state()
        for k, v in kw.items():
            setattr(s, k, v)
        self._o._set_state(s)

    @property
    def tapsPerphi(self):
        return self._o._taps_per_phase

    @property
    def Nphi(self):
        return self._o._n_phi


class _HasDeficit:
    @property
    def inputDeficit(self):
        return self._st().input_deficit

    @inputDeficit.setter
    def inputDeficit(self, v):          # examples/FIRFarrow.jl:29 pokes this field directly
        self._set(input_deficit=int(v))


class _HasPfb:
    @property
    def pfb(self):
        return self._o._pfb(0)


class FIRStandard(FIRKernel):
    @property
    def h(self):
        return self._o._pfb(0)[:, 0]

    @property
    def hLen(self):
        return self._o._h_len


class FIRInterpolator(FIRKernel, _HasPfb):
    @property
    def interpolation(self):
        return self._o._ratio.numerator


class FIRDecimator(FIRKernel, _HasDeficit):
    @property
    def h(self):
        return self._o._pfb(0)[:, 0]

    @property
    def hLen(self):
        return self._o._h_len

    @property
    def decimation(self):
        return self._o._ratio.denominator


class FIRRational(FIRKernel, _HasDeficit, _HasPfb):
    @property
    def ratio(self):
        return self._o._ratio

    @property
    def phiIdx(self):
        return self._st().phi_idx

    @phiIdx.setter
    def phiIdx(self, v):
        self._set(phi_idx=int(v))


class FIRArbitrary(FIRKernel, _HasDeficit, _HasPfb):
    @property
    def rate(self):
        return self._o._rate

    @property
    def dpfb(self):
        return self._o._pfb(1)

    @property
    def phiAccumulator(self):
        return self._st().phi_accumulator

    @property
    def phiIdx(self):
        return self._st().phi_idx

    @property
    def alpha(self):
        return self._st().alpha

    @property
    def delta(self):
        return self._o._n_phi / self._o._rate

    @property
    def xIdx(self):
        return self._st().x_idx


class FIRFarrow(FIRKernel, _HasDeficit, _HasPfb):
    @property
    def rate(self):
        return self._o._rate

    @property
    def pnfb(self):
        return self._o._pnfb

    @property
    def polyorder(self):
        return self._o._polyorder

    @property
    def phiIdx(self):
        return self._st().phi_accumulator       # Float64 phase, src/Filters.jl:131

    @property
    def delta(self):
        return self._o._n_phi / self._o._rate

    @property
    def xIdx(self):
        return self._st().x_idx

    @property
    def currentTaps(self):
        return tapsforphase(self, self.phiIdx)


_KERNEL_CLASS = {_ffi.STANDARD: FIRStandard, _ffi.INTERPOLATOR: FIRInterpolator, _ffi.DECIMATOR: FIRDecimator,
                 _ffi.RATIONAL: FIRRational, _ffi.ARBITRARY: FIRArbitrary, _ffi.FARROW: FIRFarrow}


def _is_torch(x):
    return type(x).__module__.startswith("torch")


# --------------------------------------------------------------------------------------
# FIRFilter
# --------------------------------------------------------------------------------------
class FIRFilter:
    """FIRFilter(h, ratio=1//1)  |  FIRFilter(h, rate::float, Nphi=32)  |  FIRFilter(h, rate::float, Nphi, polyorder)
    (src/Filters.jl:158-198).  The kernel type is chosen from the ratio exactly as upstream.

    The device handle is created at the first `filt` (sample dtype and channel count come from x, as the
    reference fixes the history eltype at first use, src/Filters.jl:452), or immediately when
    `sample_dtype` and `nchannels` are given.  `device=-1` makes a host-only handle (sequencing and
    state calls only; filt raises)."""

    def __init__(self, h, ratio=Fraction(1, 1), Nphi=None, polyorder=None, *, nchannels=None, sample_dtype=None,
                 device=0, pnfb=None):
        h = np.ascontiguousarray(h)
        if h.dtype not in (np.float32, np.float64):
            h = h.astype(np.float64)
        if h.ndim != 1 or len(h) < 1:
            raise ValueError("h must be a non-empty vector")
        self._h = h
        self._h_len = len(h)
        self._device = device
        self._handle = None
        self._pnfb = None
        self._polyorder = -1
        if isinstance(ratio, (float, np.floating)):
            self._rate = float(ratio)
            if not self._rate > 0.0:
                raise ValueError("rate must be greater than 0")              # src/Filters.jl:184,193
            self._ratio = None
            self._n_phi = 32 if Nphi is None else int(Nphi)
            self._taps_per_phase = -(-len(h) // self._n_phi)
            if polyorder is None:
                self._kind = _ffi.ARBITRARY
            else:
                self._kind = _ffi.FARROW
                self._polyorder = int(polyorder)
                # Farrow coefficients: the library's fit (mrb_pfb2pnfb), or the caller's own as data
                self._pnfb = (np.ascontiguousarray(pfb2pnfb(taps2pfb(h, self._n_phi), self._polyorder)) if pnfb is None else
                              np.ascontiguousarray(np.asarray(pnfb, dtype=np.float64).reshape(self._taps_per_phase, self._polyorder + 1)))
        else:
            self._rate = 0.0
            self._ratio = Fraction(ratio)
            if self._ratio <= 0:
                raise ValueError("resampling ratio must be positive")
            L, M = self._ratio.numerator, self._ratio.denominator
            self._kind = (_ffi.STANDARD if (L == 1 and M == 1) else _ffi.DECIMATOR if L == 1
                          else _ffi.INTERPOLATOR if M == 1 else _ffi.RATIONAL)
            self._n_phi = 1 if self._kind in (_ffi.STANDARD, _ffi.DECIMATOR) else L
            self._taps_per_phase = len(h) if self._n_phi == 1 else -(-len(h) // L)
        self.historyLen = self._taps_per_phase - 1                           # :165,168,171,174,186,195
        self.kernel = _KERNEL_CLASS[self._kind](self)
        # state held host-side until a handle exists (so kernel fields can be poked before the first filt)
        self._pending_state = None
        self._nch = None
        self._tx = None
        if nchannels is not None and sample_dtype is not None:
            self._ensure(np.dtype(sample_dtype), int(nchannels))

    def _ctor_args(self):
        """Positional arguments that rebuild this filter (a probe or a per-segment twin of it)."""
        if self._ratio is not None:
            return (self._h, self._ratio)
        return (self._h, self._rate, self._n_phi) + ((self._polyorder,) if self._kind == _ffi.FARROW else ())

    # ---- handle management ------------------------------------------------
    def _ensure(self, tx, nch):
        tx = np.dtype(tx)
        if self._handle is not None:
            if tx != self._tx or nch != self._nch:
                raise ValueError("this FIRFilter was bound to %d channel(s) of %s; got %d of %s"
                                 % (self._nch, self._tx, nch, tx))
            return
        if tx not in _DT:
            raise TypeError("unsupported sample dtype %s" % tx)
        d = _ffi.Desc()
        d.kind = self._kind
        d.tap_dtype = _DT[self._h.dtype]
        d.sample_dtype = _DT[tx]
        d.device = self._device
        d.h = self._h.ctypes.data
        d.h_len = len(self._h)
        d.interpolation = self._ratio.numerator if self._ratio is not None else 1
        d.decimation = self._ratio.denominator if self._ratio is not None else 1
        d.rate = self._rate
        d.n_phi = self._n_phi
        d.poly_order = self._polyorder
        d.poly_coeffs = self._pnfb.ctypes.data if self._pnfb is not None else None
        d.n_channels = nch
        hd = C.c_void_p()
        _ffi.check(_ffi.lib().mrb_create(C.byref(d), C.byref(hd)))
        self._handle, self._tx, self._nch = hd, tx, nch
        self._ty = np.result_type(self._h.dtype, tx)
        if self._pending_state is not None:
            s, self._pending_state = self._pending_state, None
            self._set_state(s)

    def _host_handle(self):
        """A handle for sequencing-only calls made before the first filt."""
        if self._handle is None:
            dev, self._device = self._device, -1
            try:
                self._ensure(np.float32, 1)
            finally:
                self._device = dev
            self._host_only = True
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None and _ffi is not None:
            try:
                _ffi.lib().mrb_destroy(h)
            except Exception:
                pass

    def _rebind_if_host_only(self, tx, nch):
        if getattr(self, "_host_only", False):
            s = self._get_state()
            _ffi.lib().mrb_destroy(self._handle)
            self._handle, self._host_only = None, False
            self._pending_state = s
        self._ensure(tx, nch)

    def _get_state(self):
        s = _ffi.State()
        _ffi.check(_ffi.lib().mrb_get_state(self._host_handle(), C.byref(s)))
        return s

    def _set_state(self, s):
        _ffi.check(_ffi.lib().mrb_set_state(self._host_handle(), C.byref(s)))

    def _pfb(self, which):
        out = np.empty((self._n_phi, self._taps_per_phase), dtype=self._h.dtype)
        _ffi.check(_ffi.lib().mrb_get_pfb(self._host_handle(), which, out.ctypes.data))
        return out.T

    @property
    def history(self):
        if self._handle is None or getattr(self, "_host_only", False):
            return np.zeros(self.historyLen)                                  # src/Filters.jl:177
        out = np.empty((self._nch, self.historyLen), dtype=self._tx)
        if out.size:
            _ffi.check(_ffi.lib().mrb_get_history(self._handle, out.ctypes.data))
        return out[0] if self._nch == 1 else out

    @history.setter
    def history(self, v):
        v = np.ascontiguousarray(v, dtype=self._tx).reshape(self._nch, self.historyLen)
        if v.size:
            _ffi.check(_ffi.lib().mrb_set_history(self._handle, v.ctypes.data))

    @property
    def launch_count(self):
        n = C.c_int64()
        _ffi.check(_ffi.lib().mrb_launch_count(self._handle, C.byref(n)))
        return n.value

    @property
    def last_kernel(self):
        return _ffi.lib().mrb_last_kernel(self._handle).decode()

    def set_timing(self, on=True):
        _ffi.check(_ffi.lib().mrb_set_timing(self._handle, int(bool(on))))

    def kernel_ms(self):
        """Mean CUDA-event duration (ms) of the filter kernel(s) per filt call since the last query."""
        ms, n = C.c_double(), C.c_int64()
        _ffi.check(_ffi.lib().mrb_get_timing(self._handle, C.byref(ms), C.byref(n)))
        return ms.value if n.value else None

    def set_host_pipeline(self, block_mib=0, n_streams=0):
        """Shape of the host-buffer pipeline of filt(numpy): input MiB per channel block and streams (0 = default)."""
        _ffi.check(_ffi.lib().mrb_set_host_pipeline(self._handle, int(block_mib), int(n_streams)))

    def set_kernel_policy(self, policy):
        _ffi.check(_ffi.lib().mrb_set_kernel_policy(self._handle, int(policy)))

    # ---- filt -----------------------------------------------------------------
    def _exact_count(self, n_in):
        n = C.c_int64()
        _ffi.check(_ffi.lib().mrb_output_count(self._host_handle(), int(n_in), C.byref(n)))
        return n.value

    def _filt_into(self, x, buffer):
        """Core of filt / filt!: returns (buffer_or_new_array, count)."""
        L = _ffi.lib()
        if _is_torch(x):
            import torch
            if not x.is_cuda:
                raise ValueError("torch input must be a CUDA tensor (use numpy for host data)")
            squeeze = x.dim() == 1
            x2 = x.unsqueeze(0) if squeeze else x
            if x2.shape[-1] > 1 and x2.stride(-1) != 1:
                x2 = x2.contiguous()
            nch, n_in = x2.shape
            tx = np.dtype(str(x2.dtype).replace("torch.", ""))
            self._rebind_if_host_only(tx, nch)
            N = self._exact_count(n_in)
            if buffer is None:
                ty = getattr(torch, str(self._ty))
                if squeeze or nch == 1:
                    buffer = torch.empty((nch, N) if not squeeze else (N,), dtype=ty, device=x.device)
                else:
                    # row pitch padded to 16 bytes: the TMA fast paths need aligned channel rows; the caller gets the
                    # (nch, N) view
                    al = max(1, 16 // np.dtype(self._ty).itemsize)
                    buffer = torch.empty((nch, (N + al - 1) // al * al), dtype=ty, device=x.device)[:, :N]
            b2 = buffer.unsqueeze(0) if buffer.dim() == 1 else buffer
            if str(b2.dtype).replace("torch.", "") != str(self._ty) or not b2.is_cuda or (b2.shape[-1] > 1 and b2.stride(-1) != 1):
                raise TypeError("buffer must be a CUDA tensor of dtype %s, time contiguous" % self._ty)
            if b2.shape[0] != nch:
                raise ValueError("buffer must have one row per channel")
            ldx = x2.stride(0) if nch > 1 else max(n_in, 1)
            ldy = b2.stride(0) if nch > 1 else max(b2.shape[1], 1)
            n_out = C.c_int64()
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _ffi.check(L.mrb_filt(self._handle, x2.data_ptr(), ldx, n_in, b2.data_ptr(), ldy, b2.shape[1],
                                  C.byref(n_out), stream))
            return buffer, n_out.value
        x = np.asarray(x)
        if x.dtype not in _DT:
            raise TypeError("unsupported sample dtype %s" % x.dtype)
        squeeze = x.ndim == 1
        x2 = x[None, :] if squeeze else x
        # time-contiguous rows with any row pitch go through as they are (a column block of a pinned matrix stays
        # pinned); anything else is copied
        if x2.ndim != 2 or (x2.shape[1] > 1 and x2.strides[1] != x2.itemsize) or \
                (x2.shape[0] > 1 and (x2.strides[0] % x2.itemsize or x2.strides[0] < x2.shape[1] * x2.itemsize)):
            x2 = np.ascontiguousarray(x2)
        nch, n_in = x2.shape
        ldx = x2.strides[0] // x2.itemsize if nch > 1 else max(n_in, 1)
        self._rebind_if_host_only(x2.dtype, nch)
        N = self._exact_count(n_in)
        if buffer is None:
            buffer = np.empty((nch, N) if not squeeze else (N,), dtype=self._ty)
        b2 = buffer[None, :] if buffer.ndim == 1 else buffer
        if b2.dtype != self._ty or (b2.shape[-1] > 1 and b2.strides[-1] != b2.itemsize) or b2.shape[0] != nch:
            raise TypeError("buffer must be a %s array with one time-contiguous row per channel" % self._ty)
        ldy = b2.strides[0] // b2.itemsize if nch > 1 else max(b2.shape[1], 1)
        n_out = C.c_int64()
        _ffi.check(L.mrb_filt_host(self._handle, x2.ctypes.data, max(ldx, 1), n_in, b2.ctypes.data, ldy, b2.shape[1],
                                   C.byref(n_out)))
        return buffer, n_out.value

    def filt(self, x):
        """filt(self, x): src/Filters.jl:475-478, 519-525, 577-587, 633-650, 744-752, 838-846.
        Returns promote_type(Th, Tx) samples; may be EMPTY for decimating kernels (README.md:53)."""
        y, _ = self._filt_into(x, None)
        return y

    def filt_(self, buffer, x):
        """filt!(buffer, self, x): returns the buffer for FIRStandard / FIRInterpolator (:472,516) and the
        number of samples written for the others (:574,630,741,835)."""
        buf, n = self._filt_into(x, buffer)
        return buf if self._kind in (_ffi.STANDARD, _ffi.INTERPOLATOR) else n

    def reset(self):
        if self._handle is not None:
            _ffi.check(_ffi.lib().mrb_reset(self._handle))
        self._pending_state = None
        return self

    def set_taps(self, h, stream=None):
        """Replace the taps in place (same length and dtype), keeping phase state and history -- an adaptive filter whose
        taps change between chunks (upstream would rebuild the FIRFilter and lose its state).
        Asynchronous: ordered on `stream` (a cudaStream_t value; default: torch's current stream when torch is loaded,
        else the stream of the handle's last call), no device synchronisation; banks are rebuilt on the device."""
        h = np.ascontiguousarray(h, dtype=self._h.dtype)
        if h.shape != self._h.shape:
            raise ValueError("set_taps keeps the tap count (%d)" % len(self._h))
        self._h = h
        if self._kind == _ffi.FARROW:
            self._pnfb = np.ascontiguousarray(pfb2pnfb(taps2pfb(h, self._n_phi), self._polyorder))
        if self._handle is not None:
            pn = self._pnfb.ctypes.data if self._pnfb is not None else None
            if stream is None and not getattr(self, "_host_only", False):
                import sys
                torch = sys.modules.get("torch")
                if torch is not None and torch.cuda.is_available():
                    stream = torch.cuda.current_stream(self._device).cuda_stream
            if stream is None:
                _ffi.check(_ffi.lib().mrb_set_taps(self._handle, h.ctypes.data, len(h), pn))
            else:
                _ffi.check(_ffi.lib().mrb_set_taps_async(self._handle, h.ctypes.data, len(h), pn, stream))
        return self

    def seek(self, n0, halo=None):
        """Long-stream segment start (no reference counterpart): put the filter in the state it
        would have after consuming `n0` samples since construction, with `halo` = the historyLen samples preceding n0
        ([nchannels, historyLen] CUDA tensor, None = zeros) as its history.  Integer ratios use the closed form,
        arbitrary / Farrow the exact host replay of the phase recurrence.  Returns the absolute index of the segment's
        first output.  The filter must be bound (nchannels=, sample_dtype= at construction, or a previous filt)."""
        if self._handle is None or getattr(self, "_host_only", False):
            if halo is not None:
                raise ValueError("seek with a halo needs a filter bound to its channels and sample dtype")
            h = self._host_handle()
        else:
            h = self._handle
        k0 = C.c_int64()
        ptr, ld, stream = None, max(self.historyLen, 1), None
        if halo is not None and self.historyLen > 0:
            import torch
            if tuple(halo.shape) != (self._nch, self.historyLen) or not halo.is_cuda or not halo.is_contiguous():
                raise ValueError("halo must be a contiguous CUDA tensor of shape (nchannels, historyLen)")
            if str(halo.dtype).replace("torch.", "") != str(self._tx):
                raise TypeError("halo dtype does not match the filter's sample dtype")
            ptr, stream = halo.data_ptr(), torch.cuda.current_stream(halo.device).cuda_stream
        _ffi.check(_ffi.lib().mrb_seek(h, int(n0), ptr, ld, C.byref(k0), stream))
        return k0.value

    def setphase(self, phi):
        if not (0 <= phi <= 1):
            raise AssertionError("phase must be in [0, 1]")                     # @assert :211,217,225
        _ffi.check(_ffi.lib().mrb_setphase(self._host_handle(), float(phi)))
        s = self._get_state()
        if self._kind == _ffi.ARBITRARY:
            return s.phi_idx, s.alpha                                           # :221
        if self._kind == _ffi.FARROW:
            return s.phi_accumulator                                            # :228
        return s.phi_idx                                                        # :213

    def outputlength(self, inputlength):
        n = C.c_int64()
        _ffi.check(_ffi.lib().mrb_outputlength(self._host_handle(), int(inputlength), C.byref(n)))
        return n.value

    def inputlength(self, outputlength):
        """inputlength(self, outputlength): the evident intent of src/Filters.jl:403-422."""
        k = self.kernel
        if self._kind == _ffi.STANDARD:
            return int(outputlength)
        if self._kind == _ffi.INTERPOLATOR:
            return inputlength(outputlength, Fraction(k.interpolation, 1), 1)
        if self._kind == _ffi.DECIMATOR:
            return inputlength(outputlength, Fraction(1, k.decimation), 1) + k.inputDeficit - 1
        if self._kind == _ffi.RATIONAL:
            return inputlength(outputlength, k.ratio, k.phiIdx) + k.inputDeficit - 1
        raise TypeError("inputlength is not defined for arbitrary-rate kernels")


# --------------------------------------------------------------------------------------
# free functions, as exported by src/Multirate.jl:26-41
# --------------------------------------------------------------------------------------
def filt(a, x, *args, **kw):
    """filt(self::FIRFilter, x)  or the one-shot forms  filt(h, x, ratio=1//1) / filt(h, x, rate, Nphi=32) /
    filt(h, x, rate, Nphi, polyorder)  (src/Filters.jl:858-873)."""
    if isinstance(a, FIRFilter):
        return a.filt(x)
    return _oneshot_filter(a, x, args, kw).filt(x)


# One-shot calls build a filter, use it once and drop it: with a GPU behind it that is device allocations, a stream and the
# plan tables -- 2-7 ms around a 0.03 ms kernel at the README's benchmark shape.  The handles of recent one-shot calls are
# kept (per thread: a handle is not thread safe) and RESET -- mrb_reset is a full re-initialisation -- when the same taps,
# ratio and input layout come again, so a repeated one-shot costs its copies and one launch.
_ONESHOT_KEEP = 8
_oneshot_tls = None


def clear_oneshot_cache():
    """Drop the handles kept for repeated one-shot filt(h, x, ...) calls (this thread's)."""
    if _oneshot_tls is not None and getattr(_oneshot_tls, "cache", None):
        _oneshot_tls.cache.clear()


def _oneshot_filter(h, x, args, kw):
    global _oneshot_tls
    import threading
    if _oneshot_tls is None:
        _oneshot_tls = threading.local()
    cache = getattr(_oneshot_tls, "cache", None)
    if cache is None:
        cache = _oneshot_tls.cache = {}
    ha = np.ascontiguousarray(h)
    try:
        dev = ("host",) if isinstance(x, np.ndarray) or not hasattr(x, "device") else ("cuda", x.device.index)
        shape = tuple(x.shape[:-1]) if getattr(x, "ndim", 1) > 1 else ()
        key = (ha.tobytes(), str(ha.dtype), tuple(args), tuple(sorted(kw.items())), str(getattr(x, "dtype", None)), shape, dev)
        hash(key)
    except Exception:
        return FIRFilter(h, *args, **kw)
    f = cache.pop(key, None)
    if f is None:
        f = FIRFilter(h, *args, **kw)
    else:
        f.reset()
    cache[key] = f                                                   # (re)inserted last: the dict is the LRU order
    while len(cache) > _ONESHOT_KEEP:
        cache.pop(next(iter(cache)))
    return f


def filt_(buffer, self, x):
    """filt!(buffer, self, x)."""
    return self.filt_(buffer, x)


def reset(self):
    return self.reset()


def setphase(self, phi):
    return (self._o if isinstance(self, FIRKernel) else self).setphase(phi)


def outputlength(a, b, initialphi=None):
    """outputlength(self, inputlength)  or  outputlength(inputlength, ratio, initialphi) (src/Filters.jl:352-385)."""
    if isinstance(a, FIRFilter):
        return a.outputlength(b)
    if isinstance(a, FIRKernel):
        return a._o.outputlength(b)
    ratio = Fraction(b)
    return int(np.ceil(((int(a) * ratio.numerator) - initialphi + 1) / ratio.denominator))


def inputlength(a, b, initialphi=None):
    """inputlength(outputlength, ratio, initialphi) (src/Filters.jl:396-401)  or  inputlength(self, outputlength)."""
    if isinstance(a, FIRFilter):
        return a.inputlength(b)
    ratio = Fraction(b)
    n = C.c_int64()
    _ffi.check(_ffi.lib().mrb_inputlength(int(a), ratio.numerator, ratio.denominator, int(initialphi), C.byref(n)))
    return n.value


def tapsforphase_(buffer, kernel, phase):
    """tapsforphase!(buffer, kernel, phase): src/Filters.jl:677-688 (arbitrary), 764-773 (farrow)."""
    o = kernel._o if isinstance(kernel, FIRKernel) else kernel
    if o._kind not in (_ffi.ARBITRARY, _ffi.FARROW):
        raise TypeError("tapsforphase is defined for FIRArbitrary and FIRFarrow kernels")
    if not (0 <= phase <= o._n_phi + 1):
        raise ValueError("phase must be >= 0 and <= Nphi+1")                    # :678,:765
    if len(buffer) < o._taps_per_phase:
        raise ValueError("buffer is too small")                                 # :679,:766
    tmp = np.empty(o._taps_per_phase, dtype=o._h.dtype)
    _ffi.check(_ffi.lib().mrb_tapsforphase(o._host_handle(), float(phase), tmp.ctypes.data))
    buffer[:o._taps_per_phase] = tmp
    return buffer


def tapsforphase(kernel, phase):
    o = kernel._o if isinstance(kernel, FIRKernel) else kernel
    return tapsforphase_(np.empty(o._taps_per_phase, dtype=o._h.dtype), kernel, phase)
