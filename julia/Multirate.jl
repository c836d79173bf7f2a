# Multirate.jl -- drop-in host side for JayKickliter/Multirate.jl's streaming polyphase FIR path, bound to libmrb
# (include/mrb.h, sm_100a) with `ccall`.  Same module name, exported names and call shapes as the reference
# (src/Multirate.jl:10-41), written in current Julia (the reference is Julia-0.3 syntax and no longer parses).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no julia binary.  tests/test_julia_binding.py checks
# every `ccall` below (symbol, arity, argument widths) against include/mrb.h; the same C-ABI is exercised end to end by
# the Python twin (multirate.jl_b200/filters.py).  See INTEGRATION.md.
#
#   using Multirate
#   h  = firdes(24*147, 0.5/147, kaiser, beta = 7.8562)
#   f  = FIRFilter(h, 147//160)            # picks FIRRational, as src/Filters.jl:158-180
#   y1 = filt(f, x1); y2 = filt(f, x2)     # state (history, phase, deficit) carried on the device
#   Y  = filt(f, X)                        # X::Matrix (nsamples x nchannels): channels in columns (additive)
#   filt!(Yd, f, Xd)                       # Xd, Yd::DeviceMatrix: device pointers, nothing crosses PCIe
#   f.kernel.inputDeficit += 3             # kernel fields are live views of the handle (examples/FIRFarrow.jl:29)
module Multirate

export  hanning, hamming, kaiser, blackman                                     # src/Multirate.jl:10-13 (typo `hammming` fixed)
export  firdes, kaiserlength, firprototype, FIRResponse, LOWPASS, HIGHPASS, BANDPASS, BANDSTOP   # :16-23 (`HIGPASS` fixed)
export  FIRFilter, FIRInterpolator, FIRArbitrary, FIRDecimator, FIRFarrow, FIRRational, FIRStandard,
        filt!, filt, setphase, tapsforphase!, tapsforphase, taps2pfb, reset, outputlength, inputlength   # :26-41
export  DeviceMatrix, seek!, schedule, settaps!                                # additive: device pointers, long streams, live taps

const libmrb = get(ENV, "LIBMRB", joinpath(@__DIR__, "..", "multirate.jl_b200", "csrc", "libmrb.so"))

# ---- enums / structs of include/mrb.h ----------------------------------------------------------------------------
const MRB_KIND_AUTO, MRB_STANDARD, MRB_INTERPOLATOR, MRB_DECIMATOR, MRB_RATIONAL, MRB_ARBITRARY, MRB_FARROW =
    Int32(-1), Int32(0), Int32(1), Int32(2), Int32(3), Int32(4), Int32(5)
dtypecode(::Type{Float32}) = Int32(0); dtypecode(::Type{Float64}) = Int32(1)
dtypecode(::Type{ComplexF32}) = Int32(2); dtypecode(::Type{ComplexF64}) = Int32(3)

struct MrbDesc
    kind::Int32; tap_dtype::Int32; sample_dtype::Int32; device::Int32
    h::Ptr{Cvoid}; h_len::Int64
    interpolation::Int64; decimation::Int64
    rate::Float64; n_phi::Int32; poly_order::Int32
    poly_coeffs::Ptr{Float64}
    n_channels::Int64
end

mutable struct MrbState
    phi_idx::Int64; input_deficit::Int64; x_idx::Int64; phi_accumulator::Float64; alpha::Float64
    MrbState() = new(0, 0, 0, 0.0, 0.0)
end

lasterror() = unsafe_string(ccall((:mrb_last_error, libmrb), Cstring, ()))
check(rc::Int32) = rc == 0 ? nothing : error(lasterror())      # reference wording, e.g. "buffer is too small"

# ---- tap design: src/FIRDesign.jl:7-95 (host side, taps are an input of the path) ---------------------------------
@enum FIRResponse LOWPASS BANDPASS HIGHPASS BANDSTOP

# windows: the reference takes them from DSP.jl (`using DSP.Windows`, src/Multirate.jl:9), which is not a dependency
# here.  kaiser(n, beta) takes the textbook Kaiser beta -- the value kaiserlength returns.
function besseli0(x::Float64)
    s, t, k = 1.0, 1.0, 1
    while true
        t *= (x / (2k))^2
        s += t
        t < 1e-17 * s && return s
        k += 1
    end
end
kaiser(n::Integer, beta::Real) = n == 1 ? [1.0] :
    [besseli0(Float64(beta) * sqrt(1 - (2k / (n - 1) - 1)^2)) / besseli0(Float64(beta)) for k in 0:n-1]
hanning(n::Integer) = [0.5 - 0.5cos(2pi * k / (n - 1)) for k in 0:n-1]
hamming(n::Integer) = [0.54 - 0.46cos(2pi * k / (n - 1)) for k in 0:n-1]
blackman(n::Integer) = [0.42 - 0.5cos(2pi * k / (n - 1)) + 0.08cos(4pi * k / (n - 1)) for k in 0:n-1]

function kaiserlength(transition::Real, attenuation::Real = 60; samplerate = 1.0)      # src/FIRDesign.jl:18-32
    transition = transition / samplerate
    numtaps = ceil(Int, (attenuation - 7.95) / (2 * pi * 2.285 * transition))
    beta = attenuation > 50 ? 0.1102 * (attenuation - 8.7) :
           attenuation >= 21 ? 0.5842 * (attenuation - 21)^0.4 + 0.07886 * (attenuation - 21) : 0.0
    return numtaps, beta
end

function firprototype(numtaps::Integer, F; response::FIRResponse = LOWPASS)              # src/FIRDesign.jl:49-65
    M = numtaps - 1
    if response == LOWPASS
        return [2 * F * sinc(2 * F * (n - M / 2)) for n in 0:M]
    elseif response == BANDPASS
        return [2 * (F[1] * sinc(2 * F[1] * (n - M / 2)) - F[2] * sinc(2 * F[2] * (n - M / 2))) for n in 0:M]
    elseif response == HIGHPASS
        M = isodd(M) ? M + 1 : M
        return [sinc(n - M / 2) - 2 * F * sinc(2 * F * (n - M / 2)) for n in 0:M]
    elseif response == BANDSTOP
        return [2 * (F[2] * sinc(2 * F[2] * (n - M / 2)) - F[1] * sinc(2 * F[1] * (n - M / 2))) for n in 0:M]
    end
    error("Not a valid FIR_TYPE")
end

function firdes(numtaps::Integer, cutoff, windowfunction::Function; response::FIRResponse = LOWPASS,
                samplerate = 1.0, beta = 6.75)                                            # src/FIRDesign.jl:76-86
    cutoff = cutoff ./ samplerate
    prototype = firprototype(numtaps, cutoff, response = response)
    numtaps = length(prototype)
    windowfunction === kaiser ? prototype .* kaiser(numtaps, beta) : prototype .* windowfunction(numtaps)
end

function firdes(cutoff, transitionwidth::Real, stopbandAttenuation::Real = 60; response::FIRResponse = LOWPASS,
                samplerate = 1.0)                                                         # src/FIRDesign.jl:88-95
    numtaps, beta = kaiserlength(transitionwidth, stopbandAttenuation; samplerate = samplerate)
    firdes(numtaps, cutoff, kaiser, response = response, samplerate = samplerate, beta = beta)
end

# ---- kernel types (src/Filters.jl:15,28,45,62,91,123): views of the handle ----------------------------------------
# The reference's kernels are mutable structs the examples poke directly (`kernel.inputDeficit += n`,
# examples/FIRFarrow.jl:29).  Here the state lives in the library handle; a kernel object is a typed view of its filter
# and `getproperty` / `setproperty!` read and write the handle (mrb_get_state / mrb_set_state / mrb_get_pfb).
abstract type FIRKernel end
mutable struct FIRStandard{T} <: FIRKernel;     owner::Any; end
mutable struct FIRInterpolator{T} <: FIRKernel; owner::Any; end
mutable struct FIRDecimator{T} <: FIRKernel;    owner::Any; end
mutable struct FIRRational{T} <: FIRKernel;     owner::Any; end
mutable struct FIRArbitrary{T} <: FIRKernel;    owner::Any; end
mutable struct FIRFarrow{T} <: FIRKernel;       owner::Any; end
kindcode(::Type{<:FIRStandard}) = MRB_STANDARD;         kindcode(::Type{<:FIRInterpolator}) = MRB_INTERPOLATOR
kindcode(::Type{<:FIRDecimator}) = MRB_DECIMATOR;       kindcode(::Type{<:FIRRational}) = MRB_RATIONAL
kindcode(::Type{<:FIRArbitrary}) = MRB_ARBITRARY;       kindcode(::Type{<:FIRFarrow}) = MRB_FARROW

# ---- FIRFilter (src/Filters.jl:151-155) -----------------------------------------------------------------------------
mutable struct FIRFilter{Tk<:FIRKernel}
    kernel::Tk
    historyLen::Int
    h::Vector                           # taps as given (Float32 or Float64)
    ratio::Rational{Int}
    rate::Float64
    Nϕ::Int
    polyorder::Int
    pnfb::Matrix{Float64}               # farrow: (order+1) x tapsPerϕ coefficients, from mrb_pfb2pnfb
    handle::Ptr{Cvoid}                  # mrb_filter*
    hostonly::Bool                      # the handle was made for sequencing calls before the first filt (device = -1)
    Tx::DataType
    nchannels::Int
    device::Int
end

tapeltype(f::FIRFilter) = eltype(f.h)

function newfilter(::Type{K}, h::Vector{Th}, ratio, rate, Nϕ, polyorder, pnfb, historyLen, device) where {K,Th}
    k = K{Th}(nothing)
    f = FIRFilter{K{Th}}(k, historyLen, copy(h), ratio, rate, Nϕ, polyorder, pnfb, C_NULL, false, Nothing, 0, device)
    setfield!(k, :owner, f)
    finalizer(release!, f)
    f
end

function release!(f::FIRFilter)
    if f.handle != C_NULL
        ccall((:mrb_destroy, libmrb), Int32, (Ptr{Cvoid},), f.handle)
        f.handle = C_NULL
    end
    nothing
end

# FIRFilter(h, ratio=1//1): src/Filters.jl:158-180
function FIRFilter(h::Vector{Th}, ratio::Rational = 1//1; device::Integer = 0) where {Th<:Union{Float32,Float64}}
    L, M = numerator(ratio), denominator(ratio)
    if ratio == 1
        newfilter(FIRStandard, h, ratio, 0.0, 1, -1, zeros(0, 0), length(h) - 1, device)               # :165
    elseif L == 1
        newfilter(FIRDecimator, h, ratio, 0.0, 1, -1, zeros(0, 0), length(h) - 1, device)              # :168
    elseif M == 1
        newfilter(FIRInterpolator, h, ratio, 0.0, L, -1, zeros(0, 0), cld(length(h), L) - 1, device)   # :171
    else
        newfilter(FIRRational, h, ratio, 0.0, L, -1, zeros(0, 0), cld(length(h), L) - 1, device)       # :174
    end
end

# FIRFilter(h, rate, Nϕ=32): src/Filters.jl:183-189
function FIRFilter(h::Vector{Th}, rate::AbstractFloat, Nϕ::Integer = 32; device::Integer = 0) where {Th<:Union{Float32,Float64}}
    rate > 0.0 || error("rate must be greater than 0")
    newfilter(FIRArbitrary, h, 1//1, Float64(rate), Int(Nϕ), -1, zeros(0, 0), cld(length(h), Nϕ) - 1, device)
end

# FIRFilter(h, rate, Nϕ, polyorder): src/Filters.jl:192-198.  The polynomial fit (pfb2pnfb, src/Filters.jl:311-321;
# polyfit, src/support.jl:85-88) is ill-conditioned and solver dependent, so every binding takes it from the library:
# mrb_pfb2pnfb is the one agreed recipe (Householder QR in Float64, coefficients rounded to the tap type).
function FIRFilter(h::Vector{Th}, rate::AbstractFloat, Nϕ::Integer, polyorder::Integer; device::Integer = 0) where {Th<:Union{Float32,Float64}}
    rate > 0.0 || error("rate must be greater than 0")
    T = cld(length(h), Nϕ)
    pnfb = Matrix{Float64}(undef, polyorder + 1, T)            # column i = coefficients of tap row i, lowest order first
    check(ccall((:mrb_pfb2pnfb, libmrb), Int32, (Ptr{Cvoid}, Int64, Int32, Int64, Int32, Ptr{Float64}),
                h, length(h), dtypecode(Th), Nϕ, polyorder, pnfb))
    newfilter(FIRFarrow, h, 1//1, Float64(rate), Int(Nϕ), Int(polyorder), pnfb, T - 1, device)
end

function create(f::FIRFilter{Tk}, ::Type{Tx}, nch::Integer, device::Integer) where {Tk,Tx}
    Th = tapeltype(f)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    h, pnfb = f.h, f.pnfb
    GC.@preserve h pnfb begin
        d = MrbDesc(kindcode(Tk), dtypecode(Th), dtypecode(Tx), Int32(device), pointer(h), length(h),
                    numerator(f.ratio), denominator(f.ratio), f.rate, Int32(f.Nϕ), Int32(f.polyorder),
                    isempty(pnfb) ? Ptr{Float64}(C_NULL) : pointer(pnfb), nch)
        check(ccall((:mrb_create, libmrb), Int32, (Ref{MrbDesc}, Ref{Ptr{Cvoid}}), d, out))
    end
    out[]
end

# Bind the filter to a sample type and channel count (the reference fixes the history eltype at the first filt,
# src/Filters.jl:452).  A host-only handle made earlier for outputlength / setphase / kernel-field calls is replaced
# by the real one and its state carried over -- `outputlength(f, n)` before the first `filt` must not break the filter.
function bind!(f::FIRFilter, ::Type{Tx}, nch::Integer) where {Tx}
    if f.handle != C_NULL && !f.hostonly
        (f.Tx == Tx && f.nchannels == nch) || error("FIRFilter is bound to $(f.nchannels) channel(s) of $(f.Tx)")
        return f
    end
    carried = nothing
    if f.handle != C_NULL                                   # host-only handle: carry its state
        carried = MrbState()
        check(ccall((:mrb_get_state, libmrb), Int32, (Ptr{Cvoid}, Ref{MrbState}), f.handle, carried))
        release!(f)
    end
    f.handle = create(f, Tx, nch, f.device)
    f.hostonly, f.Tx, f.nchannels = false, Tx, Int(nch)
    carried === nothing || check(ccall((:mrb_set_state, libmrb), Int32, (Ptr{Cvoid}, Ref{MrbState}), f.handle, carried))
    f
end

function hosthandle(f::FIRFilter)                            # sequencing / state calls before the first filt
    if f.handle == C_NULL
        f.handle = create(f, Float32, 1, -1)
        f.hostonly, f.Tx, f.nchannels = true, Float32, 1
    end
    f.handle
end

getstate(f::FIRFilter) = (s = MrbState(); check(ccall((:mrb_get_state, libmrb), Int32, (Ptr{Cvoid}, Ref{MrbState}), hosthandle(f), s)); s)
setstate!(f::FIRFilter, s::MrbState) = check(ccall((:mrb_set_state, libmrb), Int32, (Ptr{Cvoid}, Ref{MrbState}), hosthandle(f), s))

function bank(f::FIRFilter, which::Integer)                  # kernel.pfb / kernel.dpfb / kernel.h in the reference's layout
    Th = tapeltype(f)
    T = f.Nϕ == 1 && !(f.kernel isa Union{FIRArbitrary,FIRFarrow}) ? length(f.h) : cld(length(f.h), f.Nϕ)
    pfb = Matrix{Th}(undef, T, f.Nϕ)
    check(ccall((:mrb_get_pfb, libmrb), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), hosthandle(f), Int32(which), pfb))
    pfb
end

# kernel fields (src/Filters.jl:15-24, 28-41, 45-58, 62-80, 91-117, 123-147)
function Base.getproperty(k::FIRKernel, name::Symbol)
    name === :owner && return getfield(k, :owner)
    f = getfield(k, :owner)::FIRFilter
    name === :inputDeficit && return Int(getstate(f).input_deficit)
    name === :xIdx && return Int(getstate(f).x_idx)
    name === :ϕIdx && return k isa FIRFarrow ? getstate(f).phi_accumulator : Int(getstate(f).phi_idx)
    name === :ϕAccumulator && return getstate(f).phi_accumulator
    name === :α && return getstate(f).alpha
    name === :Nϕ && return f.Nϕ
    name === :tapsPerϕ && return cld(length(f.h), f.Nϕ)
    name === :pfb && return bank(f, 0)
    name === :dpfb && return bank(f, 1)
    name === :h && return vec(bank(f, 0))                   # flipped taps, as FIRStandard / FIRDecimator store them
    name === :hLen && return length(f.h)
    name === :ratio && return f.ratio
    name === :interpolation && return numerator(f.ratio)
    name === :decimation && return denominator(f.ratio)
    name === :criticalYidx && return fld(cld(length(f.h), f.Nϕ) * numerator(f.ratio), denominator(f.ratio))   # :77, unused upstream
    name === :rate && return f.rate
    name === :Δ && return f.Nϕ / f.rate
    name === :polyorder && return f.polyorder
    name === :pnfb && return f.pnfb
    name === :currentTaps && return tapsforphase(f, getstate(f).phi_accumulator)
    error("type $(typeof(k)) has no field $name")
end

function Base.setproperty!(k::FIRKernel, name::Symbol, v)
    f = getfield(k, :owner)::FIRFilter
    s = getstate(f)
    if name === :inputDeficit
        s.input_deficit = Int64(v)
    elseif name === :xIdx
        s.x_idx = Int64(v)
    elseif name === :ϕIdx
        k isa FIRFarrow ? (s.phi_accumulator = Float64(v)) : (s.phi_idx = Int64(v))
    elseif name === :ϕAccumulator
        s.phi_accumulator = Float64(v)
    elseif name === :α
        s.alpha = Float64(v)
    else
        error("field $name of $(typeof(k)) cannot be assigned")
    end
    setstate!(f, s)
    v
end

exactcount(f::FIRFilter, n::Integer) = (r = Ref{Int64}(0);
    check(ccall((:mrb_output_count, libmrb), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}), hosthandle(f), n, r)); r[])

returnsbuffer(::FIRFilter{<:Union{FIRStandard,FIRInterpolator}}) = true      # src/Filters.jl:472,516
returnsbuffer(::FIRFilter) = false                                            # :574,630,741,835 return the count

# ---- filt! / filt on host arrays ------------------------------------------------------------------------------------
# filt!(buffer, self, x): returns the buffer for FIRStandard / FIRInterpolator and the number of samples written
# otherwise.  Matrix arguments hold one channel per column (additive).
function filt!(buffer::VecOrMat{Tb}, f::FIRFilter, x::VecOrMat{Tx}) where {Tb,Tx}
    Tb == promote_type(tapeltype(f), Tx) || error("buffer eltype must be $(promote_type(tapeltype(f), Tx))")
    size(buffer, 2) == size(x, 2) || error("buffer must have one column per channel")
    bind!(f, Tx, size(x, 2))
    n = Ref{Int64}(0)
    check(ccall((:mrb_filt_host, libmrb), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Ref{Int64}),
                f.handle, x, max(size(x, 1), 1), size(x, 1), buffer, max(size(buffer, 1), 1), size(buffer, 1), n))
    returnsbuffer(f) ? buffer : Int(n[])
end

# filt(self, x): src/Filters.jl:475-478, 519-525, 577-587, 633-650, 744-752, 838-846 (may return an EMPTY array)
function filt(f::FIRFilter, x::Vector{Tx}) where {Tx}
    bind!(f, Tx, 1)
    y = Vector{promote_type(tapeltype(f), Tx)}(undef, exactcount(f, length(x)))
    filt!(y, f, x)
    y
end
function filt(f::FIRFilter, x::Matrix{Tx}) where {Tx}                            # channels in columns
    bind!(f, Tx, size(x, 2))
    y = Matrix{promote_type(tapeltype(f), Tx)}(undef, exactcount(f, size(x, 1)), size(x, 2))
    filt!(y, f, x)
    y
end

# ---- filt! on device memory: the streaming path of north_star (state on the device, no host round trip) -------------
# A DeviceMatrix is a plain description of device memory the caller owns (e.g. `pointer(::CuArray)` from CUDA.jl, or
# any allocator): nsamples x nchannels, column-major with leading dimension ld (samples).  No CUDA.jl kernels and no
# dependency on CUDA.jl here; the library launches on `stream` (a cudaStream_t, C_NULL = default stream).
struct DeviceMatrix{T}
    ptr::Ptr{Cvoid}
    nsamples::Int
    nchannels::Int
    ld::Int
end
DeviceMatrix{T}(ptr, nsamples::Integer, nchannels::Integer = 1) where {T} = DeviceMatrix{T}(ptr, nsamples, nchannels, max(nsamples, 1))

# filt!(Y, self, X; stream): asynchronous; returns the per-channel output count (computed on the host from the closed
# form / exact phase replay, so no device synchronisation is needed to size the next call)
function filt!(y::DeviceMatrix{Tb}, f::FIRFilter, x::DeviceMatrix{Tx}; stream::Ptr{Cvoid} = C_NULL) where {Tb,Tx}
    Tb == promote_type(tapeltype(f), Tx) || error("buffer eltype must be $(promote_type(tapeltype(f), Tx))")
    y.nchannels == x.nchannels || error("buffer must have one column per channel")
    bind!(f, Tx, x.nchannels)
    n = Ref{Int64}(0)
    check(ccall((:mrb_filt, libmrb), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Ref{Int64}, Ptr{Cvoid}),
                f.handle, x.ptr, x.ld, x.nsamples, y.ptr, y.ld, y.nsamples, n, stream))
    Int(n[])
end

# one-shot forms, src/Filters.jl:858-873.  Building a filter costs device allocations, a stream and the plan tables (milliseconds
# around a 0.03 ms kernel at the README's benchmark shape), so the filters of recent one-shot calls are kept per task and RESET
# (mrb_reset is a full re-initialisation) when the same taps / ratio / input layout come again -- the Python twin does the same
# (filters.py _oneshot_filter; measured there: 3.5 ms -> 0.56 ms per repeated call).
const ONESHOT_KEEP = 8
oneshot_cache() = get!(() -> Vector{Pair{Any,Any}}(), task_local_storage(), :multirate_oneshot)::Vector{Pair{Any,Any}}

function oneshot_filter(h::Vector, x::VecOrMat, args...)
    cache = oneshot_cache()
    key = (copy(h), args, eltype(x), size(x, 2))
    i = findfirst(p -> isequal(p.first, key), cache)
    if i === nothing
        f = FIRFilter(h, args...)
    else
        f = cache[i].second
        deleteat!(cache, i)
        reset(f)
    end
    push!(cache, key => f)                                   # most recent last
    length(cache) > ONESHOT_KEEP && popfirst!(cache)
    return f
end
clear_oneshot_cache() = empty!(oneshot_cache())

filt(h::Vector, x::VecOrMat, ratio::Rational = 1//1) = filt(oneshot_filter(h, x, ratio), x)
filt(h::Vector, x::VecOrMat, rate::AbstractFloat, Nϕ::Integer = 32) = filt(oneshot_filter(h, x, rate, Nϕ), x)
filt(h::Vector, x::VecOrMat, rate::AbstractFloat, Nϕ::Integer, polyorder::Integer) = filt(oneshot_filter(h, x, rate, Nϕ, polyorder), x)

# ---- state, lengths, utilities ----------------------------------------------------------------------------------------
# reset(self): src/Filters.jl:244-260, defined as full re-initialisation (SURVEY 9.2)
reset(f::FIRFilter) = (f.handle != C_NULL && check(ccall((:mrb_reset, libmrb), Int32, (Ptr{Cvoid},), f.handle)); f)

# setphase(self, ϕ), ϕ in [0, 1]: src/Filters.jl:210-232 (definitions per SURVEY 9.1, 9.8)
function setphase(f::FIRFilter, ϕ::Real)
    @assert 0 <= ϕ <= 1
    check(ccall((:mrb_setphase, libmrb), Int32, (Ptr{Cvoid}, Float64), hosthandle(f), Float64(ϕ)))
    s = getstate(f)
    f.kernel isa FIRArbitrary ? (Int(s.phi_idx), s.alpha) : f.kernel isa FIRFarrow ? s.phi_accumulator : Int(s.phi_idx)
end
setphase(k::FIRKernel, ϕ::Real) = setphase(getfield(k, :owner)::FIRFilter, ϕ)

# long-stream segment start (no upstream counterpart): state after n0 consumed samples; halo = the historyLen samples
# before n0 as a device pointer (C_NULL = zeros), one row of ldhalo samples per channel.  Returns the first output index.
function seek!(f::FIRFilter, n0::Integer, halo::Ptr{Cvoid} = C_NULL, ldhalo::Integer = 0, stream::Ptr{Cvoid} = C_NULL)
    k0 = Ref{Int64}(0)
    check(ccall((:mrb_seek, libmrb), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ref{Int64}, Ptr{Cvoid}),
                hosthandle(f), n0, halo, ldhalo, k0, stream)); k0[]
end

# per-output schedule of the next `n` inputs, state untouched: (0-based index of each window's last input sample,
# 0-based branch, α or Farrow phase) -- the loop variables of src/Filters.jl:558-569, 613-625, 717-732, 814-826
function schedule(f::FIRFilter, n::Integer)
    cnt = exactcount(f, n)
    idx = Vector{Int64}(undef, cnt); branch = Vector{Int32}(undef, cnt); frac = Vector{Float64}(undef, cnt)
    check(ccall((:mrb_get_schedule, libmrb), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}),
                hosthandle(f), n, idx, branch, frac))
    idx, branch, frac
end

# live tap update (no upstream counterpart): same tap count, phase state and history kept; Farrow filters are refitted
# by the library (poly_coeffs = NULL)
function settaps!(f::FIRFilter, h::Vector{Th}) where {Th<:Union{Float32,Float64}}
    Th == tapeltype(f) && length(h) == length(f.h) || error("settaps! keeps the tap count and type")
    check(ccall((:mrb_set_taps, libmrb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}),
                hosthandle(f), h, length(h), Ptr{Float64}(C_NULL)))
    f.h = copy(h)
    if f.kernel isa FIRFarrow
        check(ccall((:mrb_pfb2pnfb, libmrb), Int32, (Ptr{Cvoid}, Int64, Int32, Int64, Int32, Ptr{Float64}),
                    h, length(h), dtypecode(Th), f.Nϕ, f.polyorder, f.pnfb))
    end
    f
end

function outputlength(f::FIRFilter, inputlength::Integer)                  # src/Filters.jl:359-385
    r = Ref{Int64}(0)
    check(ccall((:mrb_outputlength, libmrb), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}), hosthandle(f), inputlength, r)); r[]
end
outputlength(inputlength::Integer, ratio::Rational, initialϕ::Integer) =    # src/Filters.jl:352-357
    ceil(Int, ((inputlength * numerator(ratio)) - initialϕ + 1) / denominator(ratio))

function inputlength(outputlength::Integer, ratio::Rational, initialϕ::Integer)     # src/Filters.jl:396-401
    r = Ref{Int64}(0)
    check(ccall((:mrb_inputlength, libmrb), Int32, (Int64, Int64, Int64, Int64, Ref{Int64}),
                outputlength, numerator(ratio), denominator(ratio), initialϕ, r)); r[]
end
# inputlength(self, outputlength): the evident intent of the dead methods at src/Filters.jl:403-422 (SURVEY 9.4)
function inputlength(f::FIRFilter, outputlength::Integer)
    k = f.kernel
    k isa FIRStandard && return Int(outputlength)
    k isa FIRInterpolator && return inputlength(outputlength, numerator(f.ratio)//1, 1)
    k isa FIRDecimator && return inputlength(outputlength, 1//denominator(f.ratio), 1) + k.inputDeficit - 1
    k isa FIRRational && return inputlength(outputlength, f.ratio, k.ϕIdx) + k.inputDeficit - 1
    error("inputlength is not defined for arbitrary-rate kernels")
end

function nextphase(currentphase::Integer, ratio::Rational)                           # src/Filters.jl:433-439
    r = Ref{Int64}(0)
    check(ccall((:mrb_nextphase, libmrb), Int32, (Int64, Int64, Int64, Ref{Int64}),
                currentphase, numerator(ratio), denominator(ratio), r)); r[]
end

function taps2pfb(h::Vector{T}, Nϕ::Integer) where {T<:Union{Float32,Float64}}     # src/Filters.jl:284-298
    pfb = Matrix{T}(undef, cld(length(h), Nϕ), Nϕ)
    check(ccall((:mrb_taps2pfb, libmrb), Int32, (Ptr{Cvoid}, Int64, Int32, Int64, Ptr{Cvoid}), h, length(h), dtypecode(T), Nϕ, pfb))
    pfb
end

# tapsforphase!(buffer, kernel, phase): src/Filters.jl:677-688 (arbitrary), 764-773 (farrow)
function tapsforphase!(buffer::Vector{T}, f::FIRFilter{<:Union{FIRArbitrary{T},FIRFarrow{T}}}, phase::Real) where {T}
    0 <= phase <= f.Nϕ + 1 || error("phase must be >= 0 and <= Nϕ+1")               # :678,765
    length(buffer) >= cld(length(f.h), f.Nϕ) || error("buffer is too small")         # :679,766
    check(ccall((:mrb_tapsforphase, libmrb), Int32, (Ptr{Cvoid}, Float64, Ptr{Cvoid}), hosthandle(f), Float64(phase), buffer))
    buffer
end
tapsforphase!(buffer::Vector, k::Union{FIRArbitrary,FIRFarrow}, phase::Real) = tapsforphase!(buffer, getfield(k, :owner)::FIRFilter, phase)
tapsforphase(f::FIRFilter, phase::Real) = tapsforphase!(Vector{tapeltype(f)}(undef, cld(length(f.h), f.Nϕ)), f, phase)
tapsforphase(k::Union{FIRArbitrary,FIRFarrow}, phase::Real) = tapsforphase(getfield(k, :owner)::FIRFilter, phase)

end # module
