# MultirateB200.jl -- drop-in host side for JayKickliter/Multirate.jl's streaming polyphase FIR path,
# bound to libmrb (include/mrb.h) with `ccall`.  Same exported names and call shapes as the reference
# (src/Multirate.jl:26-41), written in current Julia (the reference is Julia-0.3 syntax).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no julia binary.  The same C-ABI is exercised
# by the Python twin (multirate.jl_b200/filters.py) in tests/; this file is the binding a Julia maintainer
# adds.  See INTEGRATION.md.
#
#   using MultirateB200
#   f  = FIRFilter(h, 147//160)            # picks FIRRational, as src/Filters.jl:158-180
#   y1 = filt(f, x1); y2 = filt(f, x2)     # state (history, phase, deficit) carried on the device
#   Y  = filt(f, X)                        # X::Matrix (nsamples x nchannels): channels in columns
module MultirateB200

import Base: filt, filt!, reset
export FIRFilter, FIRStandard, FIRInterpolator, FIRDecimator, FIRRational, FIRArbitrary, FIRFarrow,
       filt, filt!, reset, setphase, outputlength, inputlength, taps2pfb, tapsforphase, tapsforphase!

const libmrb = get(ENV, "LIBMRB", joinpath(@__DIR__, "..", "multirate.jl_b200", "csrc", "libmrb.so"))

# ---- enums / structs of include/mrb.h --------------------------------------------------------------
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

# ---- kernel tags (src/Filters.jl:15,28,45,62,91,123) ------------------------------------------------
abstract type FIRKernel end
struct FIRStandard <: FIRKernel end
struct FIRInterpolator <: FIRKernel end
struct FIRDecimator <: FIRKernel end
struct FIRRational <: FIRKernel end
struct FIRArbitrary <: FIRKernel end
struct FIRFarrow <: FIRKernel end
const KINDS = (FIRStandard, FIRInterpolator, FIRDecimator, FIRRational, FIRArbitrary, FIRFarrow)

# ---- FIRFilter ---------------------------------------------------------------------------------------
mutable struct FIRFilter{Tk<:FIRKernel,Th}
    h::Vector{Th}
    ratio::Rational{Int}
    rate::Float64
    Nϕ::Int
    polyorder::Int
    pnfb::Matrix{Float64}               # farrow: (order+1) x tapsPerϕ, host fitted
    handle::Ptr{Cvoid}                  # mrb_filter*, created at first filt (sample type and channel count)
    Tx::DataType
    nchannels::Int
    device::Int
end

kindof(ratio::Rational) = ratio == 1 ? FIRStandard : numerator(ratio) == 1 ? FIRDecimator :
                          denominator(ratio) == 1 ? FIRInterpolator : FIRRational

# FIRFilter(h, ratio=1//1): src/Filters.jl:158-180
function FIRFilter(h::Vector{Th}, ratio::Rational=1//1; device::Integer=0) where {Th<:Union{Float32,Float64}}
    FIRFilter{kindof(ratio),Th}(copy(h), ratio, 0.0, 1, -1, zeros(0, 0), C_NULL, Nothing, 0, device)
end

# FIRFilter(h, rate, Nϕ=32): src/Filters.jl:183-189
function FIRFilter(h::Vector{Th}, rate::AbstractFloat, Nϕ::Integer=32; device::Integer=0) where {Th<:Union{Float32,Float64}}
    rate > 0.0 || error("rate must be greater than 0")
    FIRFilter{FIRArbitrary,Th}(copy(h), 1//1, Float64(rate), Nϕ, -1, zeros(0, 0), C_NULL, Nothing, 0, device)
end

# FIRFilter(h, rate, Nϕ, polyorder): src/Filters.jl:192-198.  The fit (pfb2pnfb, src/Filters.jl:311-321;
# polyfit, src/support.jl:85-88) stays on the host and crosses the ABI as data.
function FIRFilter(h::Vector{Th}, rate::AbstractFloat, Nϕ::Integer, polyorder::Integer; device::Integer=0) where {Th<:Union{Float32,Float64}}
    rate > 0.0 || error("rate must be greater than 0")
    pfb = taps2pfb(h, Nϕ)
    A = Float64[x^p for x in 1:Nϕ, p in 0:polyorder]
    pnfb = zeros(polyorder + 1, size(pfb, 1))
    for i in 1:size(pfb, 1)
        pnfb[:, i] = Float64.(Th.(A \ Float64.(pfb[i, :])))         # coefficients stored as Poly{T}
    end
    FIRFilter{FIRFarrow,Th}(copy(h), 1//1, Float64(rate), Nϕ, polyorder, pnfb, C_NULL, Nothing, 0, device)
end

function bind!(f::FIRFilter{Tk,Th}, ::Type{Tx}, nch::Integer) where {Tk,Th,Tx}
    if f.handle != C_NULL
        (f.Tx == Tx && f.nchannels == nch) || error("FIRFilter is bound to $(f.nchannels) channel(s) of $(f.Tx)")
        return f
    end
    kind = Int32(findfirst(==(Tk), KINDS) - 1)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve f begin
        d = MrbDesc(kind, dtypecode(Th), dtypecode(Tx), Int32(f.device), pointer(f.h), length(f.h),
                    numerator(f.ratio), denominator(f.ratio), f.rate, Int32(f.Nϕ), Int32(f.polyorder),
                    isempty(f.pnfb) ? Ptr{Float64}(C_NULL) : pointer(f.pnfb), nch)
        check(ccall((:mrb_create, libmrb), Int32, (Ref{MrbDesc}, Ref{Ptr{Cvoid}}), d, out))
    end
    f.handle, f.Tx, f.nchannels = out[], Tx, nch
    finalizer(x -> (x.handle != C_NULL && ccall((:mrb_destroy, libmrb), Int32, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), f)
    f
end

exactcount(f::FIRFilter, n::Integer) = (r = Ref{Int64}(0);
    check(ccall((:mrb_output_count, libmrb), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}), f.handle, n, r)); r[])

# ---- filt! / filt (host arrays; device pointers go through mrb_filt the same way) -------------------------
# filt!(buffer, self, x): returns the buffer for FIRStandard / FIRInterpolator (src/Filters.jl:472,516) and the
# number of samples written otherwise (:574,630,741,835).
function filt!(buffer::VecOrMat{Tb}, f::FIRFilter{Tk,Th}, x::VecOrMat{Tx}) where {Tb,Tk,Th,Tx}
    Tb == promote_type(Th, Tx) || error("buffer eltype must be $(promote_type(Th, Tx))")
    bind!(f, Tx, size(x, 2))
    n = Ref{Int64}(0)
    check(ccall((:mrb_filt_host, libmrb), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Ref{Int64}),
                f.handle, x, max(size(x, 1), 1), size(x, 1), buffer, max(size(buffer, 1), 1), size(buffer, 1), n))
    Tk <: Union{FIRStandard,FIRInterpolator} ? buffer : Int(n[])
end

# filt(self, x): src/Filters.jl:475-478, 519-525, 577-587, 633-650, 744-752, 838-846 (may return an EMPTY array)
function filt(f::FIRFilter{Tk,Th}, x::Vector{Tx}) where {Tk,Th,Tx}
    bind!(f, Tx, 1)
    y = Vector{promote_type(Th, Tx)}(undef, exactcount(f, length(x)))
    filt!(y, f, x)
    y
end
function filt(f::FIRFilter{Tk,Th}, x::Matrix{Tx}) where {Tk,Th,Tx}       # channels in columns
    bind!(f, Tx, size(x, 2))
    y = Matrix{promote_type(Th, Tx)}(undef, exactcount(f, size(x, 1)), size(x, 2))
    filt!(y, f, x)
    y
end

# one-shot forms, src/Filters.jl:858-873
filt(h::Vector, x::VecOrMat, ratio::Rational=1//1) = filt(FIRFilter(h, ratio), x)
filt(h::Vector, x::VecOrMat, rate::AbstractFloat, Nϕ::Integer=32) = filt(FIRFilter(h, rate, Nϕ), x)
filt(h::Vector, x::VecOrMat, rate::AbstractFloat, Nϕ::Integer, polyorder::Integer) = filt(FIRFilter(h, rate, Nϕ, polyorder), x)

# ---- state, lengths, utilities -----------------------------------------------------------------------
function hosthandle(f::FIRFilter{Tk,Th}) where {Tk,Th}        # sequencing calls before the first filt
    f.handle != C_NULL && return f.handle
    dev = f.device; f.device = -1
    try bind!(f, Float32, 1) finally f.device = dev end
    f.handle
end
reset(f::FIRFilter) = (f.handle != C_NULL && check(ccall((:mrb_reset, libmrb), Int32, (Ptr{Cvoid},), f.handle)); f)
function setphase(f::FIRFilter, ϕ::Real)
    @assert 0 <= ϕ <= 1
    check(ccall((:mrb_setphase, libmrb), Int32, (Ptr{Cvoid}, Float64), hosthandle(f), ϕ))
    s = MrbState(); check(ccall((:mrb_get_state, libmrb), Int32, (Ptr{Cvoid}, Ref{MrbState}), f.handle, s)); s
end
# long-stream segment start (no upstream counterpart): state after n0 consumed samples; halo = the historyLen samples
# before n0 as a device pointer (C_NULL = zeros), one row of ldhalo samples per channel.  Returns the first output index.
function seek!(f::FIRFilter, n0::Integer, halo::Ptr{Cvoid}=C_NULL, ldhalo::Integer=0, stream::Ptr{Cvoid}=C_NULL)
    k0 = Ref{Int64}(0)
    check(ccall((:mrb_seek, libmrb), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ref{Int64}, Ptr{Cvoid}),
                hosthandle(f), n0, halo, ldhalo, k0, stream)); k0[]
end
# per-output schedule of the next `n` inputs, state untouched: (0-based index of each window's last input sample,
# 0-based branch, α or Farrow phase) -- the loop variables of src/Filters.jl:558-569, 613-625, 717-732, 814-826
function schedule(f::FIRFilter, n::Integer)
    cnt = Ref{Int64}(0)
    check(ccall((:mrb_output_count, libmrb), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}), hosthandle(f), n, cnt))
    idx = Vector{Int64}(undef, cnt[]); branch = Vector{Int32}(undef, cnt[]); frac = Vector{Float64}(undef, cnt[])
    check(ccall((:mrb_get_schedule, libmrb), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}),
                hosthandle(f), n, idx, branch, frac))
    idx, branch, frac
end
# live tap update (no upstream counterpart): same tap count, phase state and history kept; Farrow filters pass the
# refitted coefficients (pfb2pnfb of the new taps, T x (order+1), row-major)
function settaps!(f::FIRFilter{Tk,Th}, h::Vector{Th}, polycoeffs::Union{Nothing,Vector{Float64}}=nothing) where {Tk,Th}
    check(ccall((:mrb_set_taps, libmrb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}),
                hosthandle(f), h, length(h), polycoeffs === nothing ? C_NULL : pointer(polycoeffs))); f
end
function outputlength(f::FIRFilter, inputlength::Integer)                  # src/Filters.jl:352-385
    r = Ref{Int64}(0)
    check(ccall((:mrb_outputlength, libmrb), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}), hosthandle(f), inputlength, r)); r[]
end
outputlength(inputlength::Integer, ratio::Rational, initialϕ::Integer) =
    ceil(Int, ((inputlength * numerator(ratio)) - initialϕ + 1) / denominator(ratio))
function inputlength(outputlength::Integer, ratio::Rational, initialϕ::Integer)     # src/Filters.jl:396-401
    r = Ref{Int64}(0)
    check(ccall((:mrb_inputlength, libmrb), Int32, (Int64, Int64, Int64, Int64, Ref{Int64}),
                outputlength, numerator(ratio), denominator(ratio), initialϕ, r)); r[]
end
function taps2pfb(h::Vector{T}, Nϕ::Integer) where {T<:Union{Float32,Float64}}     # src/Filters.jl:284-298
    pfb = Matrix{T}(undef, cld(length(h), Nϕ), Nϕ)
    check(ccall((:mrb_taps2pfb, libmrb), Int32, (Ptr{Cvoid}, Int64, Int32, Int64, Ptr{Cvoid}), h, length(h), dtypecode(T), Nϕ, pfb))
    pfb
end
function tapsforphase!(buffer::Vector{T}, f::FIRFilter{Tk,T}, phase::Real) where {Tk<:Union{FIRArbitrary,FIRFarrow},T}
    0 <= phase <= f.Nϕ + 1 || error("phase must be >= 0 and <= Nϕ+1")               # src/Filters.jl:678,765
    length(buffer) >= cld(length(f.h), f.Nϕ) || error("buffer is too small")         # :679,766
    check(ccall((:mrb_tapsforphase, libmrb), Int32, (Ptr{Cvoid}, Float64, Ptr{Cvoid}), hosthandle(f), phase, buffer))
    buffer
end
tapsforphase(f::FIRFilter{Tk,T}, phase::Real) where {Tk,T} = tapsforphase!(Vector{T}(undef, cld(length(f.h), f.Nϕ)), f, phase)

end # module
