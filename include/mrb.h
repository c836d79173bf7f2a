/* mrb.h -- C-ABI of libmrb: Multirate.jl's streaming polyphase FIR path on B200 (sm_100a).
 *
 * The reference (JayKickliter/Multirate.jl) has no FFI boundary; its seam is Julia
 * multiple dispatch on `filt!(buffer, self::FIRFilter{K{Th}}, x::Vector{Tx})` and
 * `filt(self, x)` (src/Filters.jl:450,475,489,519,536,577,598,633,693,744,795,838).
 * Each entry point below names the reference code it replaces (file:line into the
 * reference tree).  A Julia maintainer binds these with `ccall`; the binding is shown
 * in INTEGRATION.md and julia/MultirateB200.jl.  Python binds them with ctypes
 * (multirate.jl_b200/_ffi.py).
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns an int32 status
 *    (MRB_OK == 0); nothing throws across the ABI; mrb_last_error() gives the
 *    thread-local message of the last failing call.
 *  - samples are channel-major: channel c occupies x[c*ld_x .. c*ld_x + n_in), time
 *    contiguous (a Julia Matrix of size (n_in, n_channels) is exactly this with
 *    ld_x = n_in).  All channels share ONE state machine (phase, deficit), as if
 *    n_channels identical FIRFilter objects were fed in lock step.
 *  - a handle is not thread-safe; distinct handles are independent.
 *  - filter state (history, phase index, input deficit, accumulators) lives in the
 *    handle: history in device memory, scalars mirrored on the host, so streaming
 *    chunks never round-trip sample data through the host.
 *  - there is NO CPU fallback: compute entry points fail with MRB_ERR_NO_DEVICE /
 *    MRB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MRB_H
#define MRB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrb_filter mrb_filter;

enum mrb_status {
    MRB_OK = 0,
    MRB_ERR_BAD_ARGUMENT = 1,     /* reference: error("rate must be greater than 0") :184,193 ; @assert :211,217,225 */
    MRB_ERR_BUFFER_TOO_SMALL = 2, /* reference: error("buffer is too small") :550 ; :460 ; :503 */
    MRB_ERR_CUDA = 3,
    MRB_ERR_UNSUPPORTED = 4,
    MRB_ERR_NO_DEVICE = 5
};

/* kernel types, src/Filters.jl:15,28,45,62,91,123 */
enum mrb_kind {
    MRB_KIND_AUTO = -1, /* choose as FIRFilter(h, ratio) / FIRFilter(h, rate, Nphi[, polyorder]) do, :158-198 */
    MRB_STANDARD = 0,
    MRB_INTERPOLATOR = 1,
    MRB_DECIMATOR = 2,
    MRB_RATIONAL = 3,
    MRB_ARBITRARY = 4,
    MRB_FARROW = 5
};

enum mrb_dtype { MRB_F32 = 0, MRB_F64 = 1, MRB_C64 = 2, MRB_C128 = 3 };

/* Constructor arguments: the union of FIRFilter's three constructors (src/Filters.jl:158,183,192). */
typedef struct mrb_desc {
    int32_t kind;          /* MRB_KIND_AUTO or an explicit kind (must agree with the ratio) */
    int32_t tap_dtype;     /* Th: MRB_F32 | MRB_F64 */
    int32_t sample_dtype;  /* Tx: any mrb_dtype; output dtype is promote_type(Th, Tx) (:476,522,581) */
    int32_t device;        /* CUDA ordinal; -1 = host-only handle: sequencing/state calls work, filt fails */
    const void *h;         /* h_len taps of tap_dtype, host memory */
    int64_t h_len;
    int64_t interpolation; /* ratio = interpolation // decimation; reduced to lowest terms like Julia's Rational */
    int64_t decimation;
    double rate;           /* > 0 selects the arbitrary-rate constructors (:183,192); 0 selects the Rational one */
    int32_t n_phi;         /* arbitrary/farrow: number of polyphase branches (default 32, :183) */
    int32_t poly_order;    /* farrow: >= 0 ; arbitrary: -1 */
    const double *poly_coeffs; /* farrow: taps_per_phase*(poly_order+1) coefficients, row i =
                                  polyfit(pfb[i,:], order) lowest order first (src/Filters.jl:311-321,
                                  src/support.jl:85-88), or NULL = the library's own fit (mrb_pfb2pnfb).
                                  The fit is ill-conditioned (cond ~ 1e6..1e8) and therefore solver dependent
                                  (SVD and QR agree to ~1e-10 relative, not bit for bit): bindings should take
                                  the coefficients from mrb_pfb2pnfb -- ONE recipe for every host language --
                                  and may pass their own only to reproduce a particular reference build. */
    int64_t n_channels;
} mrb_desc;

/* Carried kernel state, 1-based exactly as the reference fields (src/Filters.jl:62-70,91-103,123-135). */
typedef struct mrb_state {
    int64_t phi_idx;         /* FIRRational.ϕIdx / FIRArbitrary.ϕIdx */
    int64_t input_deficit;   /* inputDeficit */
    int64_t x_idx;           /* xIdx (arbitrary, farrow) */
    double phi_accumulator;  /* FIRArbitrary.ϕAccumulator ; FIRFarrow.ϕIdx (Float64) */
    double alpha;            /* FIRArbitrary.α */
} mrb_state;

typedef struct mrb_info {
    int32_t kind, tap_dtype, sample_dtype, out_dtype, device;
    int32_t n_phi, poly_order;
    int64_t taps_per_phase, history_len, h_len, interpolation, decimation, n_channels;
    double rate;
} mrb_info;

/* FIRFilter(h, ratio) / FIRFilter(h, rate, Nphi) / FIRFilter(h, rate, Nphi, polyorder): src/Filters.jl:158-198 */
int32_t mrb_create(const mrb_desc *desc, mrb_filter **out);
int32_t mrb_destroy(mrb_filter *f);
int32_t mrb_get_info(const mrb_filter *f, mrb_info *info);

/* outputlength(self, inputlength): src/Filters.jl:352-385 -- exact for standard / interpolator /
 * decimator / rational, an UPPER BOUND for arbitrary / farrow (the loop decides, :749,843). */
int32_t mrb_outputlength(const mrb_filter *f, int64_t n_in, int64_t *n_out);
/* exact number of outputs the next filt of n_in samples will produce (closed form / exact replay). */
int32_t mrb_output_count(const mrb_filter *f, int64_t n_in, int64_t *n_out);
/* inputlength(outputlength, ratio, initialϕ): src/Filters.jl:396-401 */
int32_t mrb_inputlength(int64_t n_out, int64_t interpolation, int64_t decimation, int64_t initial_phi, int64_t *n_in);
/* nextphase(currentphase, ratio): src/Filters.jl:433-439 (1-based phases) */
int32_t mrb_nextphase(int64_t current_phase, int64_t interpolation, int64_t decimation, int64_t *next_phase);
/* taps2pfb(h, Nphi): src/Filters.jl:284-298 ; pfb is column-major taps_per_phase x n_phi like the Julia Matrix */
int32_t mrb_taps2pfb(const void *h, int64_t h_len, int32_t dtype, int64_t n_phi, void *pfb);

/* pfb2pnfb(taps2pfb(h, Nphi), order): src/Filters.jl:311-321 with polyfit src/support.jl:85-88 -- the agreed recipe
 * for the Farrow coefficients: per tap row i the least-squares polynomial through pfb[i, phi], phi = 1..Nphi, on the
 * Vandermonde matrix A[phi, p] = phi^p solved by Householder QR in Float64 (what Julia's `A \ y` means for a full-rank
 * rectangular A), coefficients lowest order first, rounded to the tap dtype (Poly{T}, :313) and returned as Float64:
 * coeffs[i*(order+1) + p].  Host only. */
int32_t mrb_pfb2pnfb(const void *h, int64_t h_len, int32_t dtype, int64_t n_phi, int32_t order, double *coeffs);

/* filt!(buffer, self, x): src/Filters.jl:450-473 (standard), 489-517 (interpolator), 536-575 (rational),
 * 598-631 (decimator), 693-742 (arbitrary), 795-836 (farrow), including history carry (shiftin!,
 * src/support.jl:61-80).  x, y are DEVICE pointers; y_capacity is the per-channel capacity of y in
 * samples; *n_out receives the per-channel output count (0 is normal, :543-547).  Asynchronous on
 * `stream` (a cudaStream_t; NULL = default stream); the count is computed on the host. */
int32_t mrb_filt(mrb_filter *f, const void *x, int64_t ld_x, int64_t n_in, void *y, int64_t ld_y,
                 int64_t y_capacity, int64_t *n_out, void *stream);
/* Same call with HOST pointers: stages through device buffers owned by the handle, synchronous.  Channel blocks
 * are pipelined H2D -> kernels -> D2H over several streams; every stream owns the schedule / tap-row buffers it
 * writes, so concurrently running blocks never share mutable state. */
int32_t mrb_filt_host(mrb_filter *f, const void *x, int64_t ld_x, int64_t n_in, void *y, int64_t ld_y,
                      int64_t y_capacity, int64_t *n_out);
/* Shape of that pipeline: input bytes per channel block (MiB) and number of streams (1..4); 0 keeps the default
 * (64 MiB, 2 streams, or the MRB_HOST_BLOCK_MIB / MRB_HOST_STREAMS environment variables). */
int32_t mrb_set_host_pipeline(mrb_filter *f, int32_t block_mib, int32_t n_streams);
/* Advance the state machine as if n_in samples had been filtered, without data (host only).
 * History is NOT updated.  Used to run the data-independent sequencing ahead / on host-only handles. */
int32_t mrb_advance(mrb_filter *f, int64_t n_in, int64_t *n_out);
/* The schedule the next n_in inputs will produce, state untouched: per output the 0-based index of the last input
 * sample of its window (the loops' inputIdx / xIdx minus one, src/Filters.jl:558-569, 613-625, 717-732, 814-826), the
 * 0-based polyphase branch, and alpha (arbitrary, :671-672) or the Float64 phase the taps are evaluated at (farrow,
 * :789).  Destinations may be NULL; each holds mrb_output_count(n_in) entries.  For output time bases and tests. */
int32_t mrb_get_schedule(mrb_filter *f, int64_t n_in, int64_t *n_idx, int32_t *branch, double *frac);

/* reset(self): src/Filters.jl:244-260 (defined here as full re-initialisation) */
int32_t mrb_reset(mrb_filter *f);
/* setphase(self, ϕ), ϕ in [0,1]: src/Filters.jl:210-232 (definitions per SURVEY 9.1, 9.8) */
int32_t mrb_setphase(mrb_filter *f, double phi);
int32_t mrb_get_state(const mrb_filter *f, mrb_state *s);
int32_t mrb_set_state(mrb_filter *f, const mrb_state *s);
/* history: n_channels rows of history_len samples (sample dtype), host memory */
int32_t mrb_get_history(mrb_filter *f, void *host_dst);
int32_t mrb_set_history(mrb_filter *f, const void *host_src);
/* tapsforphase!(buffer, kernel, phase): src/Filters.jl:677-688 (arbitrary), 764-773 (farrow); taps in tap dtype */
int32_t mrb_tapsforphase(const mrb_filter *f, double phase, void *taps);
/* kernel.pfb / kernel.dpfb / kernel.h read-back in the reference's layout (column-major T x Nphi, tap dtype);
 * which = 0: pfb (or flipped h), 1: dpfb */
int32_t mrb_get_pfb(const mrb_filter *f, int32_t which, void *dst);

/* Long-stream segment split (no reference counterpart).  Positions the state machine
 * as if n0 input samples had already been consumed since construction.  Integer ratios: closed form,
 * first output k0 = ceil(n0*L/M), phase (k0*M) mod L, deficit floor(k0*M/L) - n0 + 1.  Arbitrary / Farrow:
 * exact host replay of the Float64 phase recurrence (src/Filters.jl:663-673, 780-786) over n0 inputs,
 * bit-identical to having filtered them (O(outputs) host work, no data touched).  Both load the
 * history from the halo = the history_len samples preceding n0 (device pointer, per channel at
 * halo + c*ld_halo; NULL = zeros).  *k0 receives the absolute index of the segment's first output. */
int32_t mrb_seek(mrb_filter *f, int64_t n0, const void *halo, int64_t ld_halo, int64_t *k0, void *stream);

/* Live tap update (no reference counterpart -- upstream rebuilds the FIRFilter and loses its state).
 * Replaces the taps in place: h has the tap dtype and length given at creation, so taps-per-phase, history length, the
 * carried phase state and the per-channel history are all kept.  ASYNCHRONOUS and ordered on `stream`, no device-wide
 * synchronisation: the banks the kernels read from global memory are rebuilt ON THE DEVICE from the raw taps (flipud /
 * taps2pfb / [diff(h); 0], src/Filters.jl:21,36,53,73,106-108,284-298), the kernels that keep taps in their parameter
 * block get them with their next launch; work already queued on the stream still filters with the old taps.
 * Farrow: poly_coeffs as in mrb_desc (NULL = the library refits, mrb_pfb2pnfb).  h may be reused as soon as the call
 * returns.  mrb_set_taps is the same call on the stream of the handle's last asynchronous call. */
int32_t mrb_set_taps_async(mrb_filter *f, const void *h, int64_t h_len, const double *poly_coeffs, void *stream);
int32_t mrb_set_taps(mrb_filter *f, const void *h, int64_t h_len, const double *poly_coeffs);

/* number of CUDA kernels this handle has launched (bench.py's gpu_launches) */
int32_t mrb_launch_count(const mrb_filter *f, int64_t *n);
/* Kernel timing for bench.py's roofline: when on, every mrb_filt brackets its FILTER kernel(s) (not the
 * history carry) with CUDA events on the launch stream.  mrb_get_timing synchronises those events, returns the
 * mean duration per mrb_filt call in milliseconds and the number of calls averaged, and clears the record. */
int32_t mrb_set_timing(mrb_filter *f, int32_t on);
int32_t mrb_get_timing(mrb_filter *f, double *mean_ms, int64_t *n_calls);
/* select the kernel family: 0 = automatic (tensor-core and tiled kernels, k_stream and k_head_warp where applicable),
 * 1 = k_generic alone (the always-correct path the parity tests use as a second opinion),
 * 2 = automatic without the tensor-core kernels (the CUDA-core fast paths: the other arm of the K5 comparison) */
int32_t mrb_set_kernel_policy(mrb_filter *f, int32_t policy);
/* name of the kernel that computed the body of the last mrb_filt on this handle ("generic", "stream", "head",
 * "tiled_c64_t24_r12", "unit_f32_l4_r8", "decim_c64_m8", "table_f32", ...; "none" before the first call) */
const char *mrb_last_kernel(const mrb_filter *f);

const char *mrb_last_error(void);
const char *mrb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MRB_H */
