"""multirate.jl_b200 -- B200-native streaming polyphase FIR (the hot path of JayKickliter/Multirate.jl).

Contents: `csrc/` (hand-written sm_100a CUDA kernels + the C-ABI of include/mrb.h, built into
csrc/libmrb.so) and the host-side mirror of the reference's operator interface (`filters.py`).
The directory name contains a dot, so import it through the repo-root shim:

    import multirate_b200 as mr          # loads this package as module "multirate_jl_b200"
"""
from . import _ffi
from ._ffi import MrbError, build
from .sharding import LongStream, channel_shard, filt_long_stream, segment_bounds, segment_plan
from .firdesign import (BANDPASS, BANDSTOP, HIGHPASS, LOWPASS, FIRResponse, blackman, firdes, firprototype, hamming, hanning,
                        kaiser, kaiserlength)
from .filters import (FIRArbitrary, FIRDecimator, FIRFarrow, FIRFilter, FIRInterpolator, FIRKernel, FIRRational,
                      FIRStandard, clear_oneshot_cache, filt, filt_, inputlength, nextphase, outputlength, pfb2pnfb, polyfit, reset,
                      setphase, taps2pfb, tapsforphase, tapsforphase_)

__all__ = ["FIRFilter", "FIRKernel", "FIRStandard", "FIRInterpolator", "FIRDecimator", "FIRRational", "FIRArbitrary",
           "FIRFarrow", "filt", "filt_", "reset", "setphase", "outputlength", "inputlength", "taps2pfb", "tapsforphase",
           "tapsforphase_", "nextphase", "polyfit", "pfb2pnfb", "MrbError", "build", "channel_shard", "segment_bounds", "segment_plan", "filt_long_stream", "LongStream", "clear_oneshot_cache",
           "firdes", "firprototype", "kaiserlength", "FIRResponse", "LOWPASS", "BANDPASS", "HIGHPASS", "BANDSTOP", "kaiser",
           "hanning", "hamming", "blackman"]
