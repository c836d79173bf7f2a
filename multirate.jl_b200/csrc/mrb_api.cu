// mrb_api.cu -- the C-ABI of include/mrb.h: handle, host sequencing, kernel dispatch.
// Compiled for sm_100a only.  There is no CPU compute path in this library.
#include "../../include/mrb.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "mrb_kernels.cuh"
#include "mrb_seq.h"
#include "mrb_tiled.cuh"
#include "mrb_unit.cuh"
#include "mrb_decim.cuh"
#include "mrb_table.cuh"
#include "mrb_mma.cuh"

using namespace mrb;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int32_t fail(int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? MRB_ERR_NO_DEVICE \
                                                                                      : MRB_ERR_CUDA,    \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);    \
    } while (0)

static size_t dsize(int d) { return d == MRB_F32 ? 4 : d == MRB_F64 ? 8 : d == MRB_C64 ? 8 : 16; }
static bool is_complex(int d) { return d == MRB_C64 || d == MRB_C128; }
static bool is_double(int d) { return d == MRB_F64 || d == MRB_C128; }
// promote_type(Th, Tx), src/Filters.jl:476,522,581
static int promote(int th, int tx) {
    const bool dbl = th == MRB_F64 || is_double(tx);
    return is_complex(tx) ? (dbl ? MRB_C128 : MRB_C64) : (dbl ? MRB_F64 : MRB_F32);
}

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct SchedSlot {            // device copy of one table-schedule sub-chunk
    int64_t *d_n = nullptr;
    int32_t *d_phi = nullptr;
    double *d_a = nullptr;
    uint64_t tag = 0;         // (call serial, slice) the slot holds: the channel blocks of one mrb_filt_host call share it
};

// Host storage of one replayed schedule: (n, phi, alpha | phase) per output.  PINNED on device-bound handles, so the
// slices are uploaded straight from it (no staging copy on the host: the arbitrary-rate path at BASELINE configs[3] is
// bound by host time per chunk, not by the kernels).  Two stores alternate between calls; `ev` marks the last upload
// that still reads a store, and is waited for before the store is overwritten.
struct SchedStore {
    int64_t *n = nullptr;
    int32_t *phi = nullptr;
    double *a = nullptr;
    size_t cap = 0, size = 0;
    bool pinned = false;
    cudaEvent_t ev = nullptr;
    bool pending = false;

    void release() {
        if (pinned) { cudaFreeHost(n); cudaFreeHost(phi); cudaFreeHost(a); }
        else { free(n); free(phi); free(a); }
        n = nullptr; phi = nullptr; a = nullptr; cap = size = 0;
        if (ev) { cudaEventDestroy(ev); ev = nullptr; }
        pending = false;
    }
    // room for `want` entries, contents kept; false when out of memory
    bool reserve(size_t want, bool pin) {
        if (want <= cap) return true;
        const size_t ncap = want + want / 8 + 64;
        int64_t *nn = nullptr; int32_t *np = nullptr; double *na = nullptr;
        if (pin) {
            if (cudaMallocHost(&nn, ncap * sizeof(int64_t)) != cudaSuccess || cudaMallocHost(&np, ncap * sizeof(int32_t)) != cudaSuccess ||
                cudaMallocHost(&na, ncap * sizeof(double)) != cudaSuccess) { cudaFreeHost(nn); cudaFreeHost(np); cudaFreeHost(na); cudaGetLastError(); return false; }
        } else {
            nn = (int64_t *)malloc(ncap * sizeof(int64_t)); np = (int32_t *)malloc(ncap * sizeof(int32_t)); na = (double *)malloc(ncap * sizeof(double));
            if (!nn || !np || !na) { free(nn); free(np); free(na); return false; }
        }
        if (size) { memcpy(nn, n, size * sizeof(int64_t)); memcpy(np, phi, size * sizeof(int32_t)); memcpy(na, a, size * sizeof(double)); }
        if (pinned) { cudaFreeHost(n); cudaFreeHost(phi); cudaFreeHost(a); } else { free(n); free(phi); free(a); }
        n = nn; phi = np; a = na; cap = ncap; pinned = pin;
        return true;
    }
};

// Everything run_channels WRITES on the device for the arbitrary-rate kinds: the uploaded schedule slices, the Farrow tap
// table of the generic path and the per-output tap rows of the table kernel.  One context per pipeline stream of
// mrb_filt_host (context 0 also serves mrb_filt): channel blocks that run concurrently on different streams never share
// a buffer, and inside one stream the stream order keeps a slice's readers ahead of the next slice's writers.
struct TableCtx {
    SchedSlot slot[2];
    bool ready = false;
    void *d_taptab = nullptr;          // farrow: R[kSchedChunk][T]
    TabRows rows;                      // table kernel: tap rows + aligned window starts of one slice
    MmaRows mrows;                     // tensor-core kernel: tap tiles + group window starts of one slice
    // the slice's head (outputs whose windows reach the history: k_head_warp / k_generic) runs on a side stream beside
    // the main kernel, which leaves issue slots free (one CTA per SM); forked and joined with events, no host wait
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
constexpr int kMaxHostStreams = 4;

struct DeviceGuard {                   // the API leaves the caller's current device as it found it
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

struct mrb_filter;
static void free_device(mrb_filter *f);
static int32_t order_after_last(mrb_filter *f, cudaStream_t st);

struct mrb_filter {
    int kind, th, tx, ty, device;
    int64_t hLen, L, M, Nphi, T, H, nch;
    int polyorder;
    double rate, delta;
    std::vector<double> bank, dbank;   // [Nphi][T] row phi = pfb[:, phi]; exact values of the Th taps
    std::vector<double> pnfb;          // [T][order+1]
    // carried state, 1-based like the reference
    int64_t phiIdx, deficit, xIdx;
    double acc, alpha;
    // device
    void *d_bank = nullptr, *d_dbank = nullptr;   // compute real type R
    double *d_pnfb = nullptr;
    void *d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    TableCtx tctx[kMaxHostStreams];
    // host-call staging
    void *d_xs = nullptr, *d_ys = nullptr;
    size_t xs_bytes = 0, ys_bytes = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t host_streams[kMaxHostStreams - 1] = {};          // further streams of the mrb_filt_host pipeline
    // stream of the last asynchronous call: a call on ANOTHER stream first waits (on the device) for that work, so the
    // history / schedule buffers of the handle are never read and written concurrently
    cudaStream_t last_stream = nullptr;
    bool last_valid = false;
    cudaEvent_t order_ev = nullptr;
    // live tap update: pinned + device staging of the raw taps, guarded by an event
    void *h_taps_stage = nullptr, *d_taps_stage = nullptr;
    size_t taps_stage_bytes = 0;
    cudaEvent_t taps_ev = nullptr;
    bool taps_pending = false;
    // table kinds: the schedule of the call in flight.  The exact replay costs ~0.3 ms per 60 K outputs, and a caller
    // typically asks for the count (to size its buffer) right before it filters: the last replay is cached, keyed by
    // the state it started from and the input length.
    SchedStore sched[2];
    int sched_cur = 0;
    bool sched_valid = false;
    mrb_state sched_from{}, sched_end{};
    int64_t sched_n_in = -1, sched_count = 0;
    TiledPlan tiled;                   // fast-path resources (mrb_tiled.cuh)
    UnitPlan unit;                     // fast path for float32 standard / interpolator (mrb_unit.cuh)
    DecPlan decim;                     // fast path for complex64 decimators (mrb_decim.cuh)
    TabPlan table;                     // fast path for arbitrary / farrow on real samples (mrb_table.cuh)
    MmaPlan mma;                       // tensor-core path for float32 samples (mrb_mma.cuh)
    uint64_t serial = 0;               // filt calls so far (tags what the table contexts hold)
    int policy = 0;
    int host_block_mib = 0, host_streams_n = 0;   // mrb_set_host_pipeline; 0 = MRB_HOST_BLOCK_MIB / MRB_HOST_STREAMS / default
    int num_sms = 148;
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev;   // one pair per timed mrb_filt
    const char *last_kernel = "none";
    int64_t launches = 0;
    mrb_filter() = default;
    mrb_filter(const mrb_filter &) = delete;
    mrb_filter &operator=(const mrb_filter &) = delete;
    ~mrb_filter() { free_device(this); }   // every failure path of mrb_create / mrb_set_taps releases what was allocated
};

static const int64_t kSchedChunk = 1 << 16;   // outputs per table-schedule sub-chunk

static void free_device(mrb_filter *f) {
    if (f->device < 0) { for (auto &st : f->sched) st.release(); return; }
    DeviceGuard guard(f->device);
    for (auto &st : f->sched) st.release();
    cudaFree(f->d_bank); cudaFree(f->d_dbank); cudaFree(f->d_pnfb);
    cudaFree(f->d_hist[0]); cudaFree(f->d_hist[1]);
    for (auto &c : f->tctx) {
        for (auto &s : c.slot) {
            cudaFree(s.d_n); cudaFree(s.d_phi); cudaFree(s.d_a);
            s = SchedSlot{};
        }
        cudaFree(c.d_taptab); c.d_taptab = nullptr;
        tabrows_release(c.rows);
        mmarows_release(c.mrows);
        if (c.side) { cudaStreamDestroy(c.side); c.side = nullptr; }
        if (c.ev_fork) { cudaEventDestroy(c.ev_fork); c.ev_fork = nullptr; }
        if (c.ev_join) { cudaEventDestroy(c.ev_join); c.ev_join = nullptr; }
        c.ready = false;
    }
    cudaFree(f->d_xs); cudaFree(f->d_ys);
    if (f->order_ev) cudaEventDestroy(f->order_ev);
    if (f->taps_ev) cudaEventDestroy(f->taps_ev);
    cudaFreeHost(f->h_taps_stage); cudaFree(f->d_taps_stage);
    tiled_release(f->tiled);
    unit_release(f->unit);
    decim_release(f->decim);
    table_release(f->table);
    mma_release(f->mma);
    for (auto &p : f->tev) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (f->own_stream) cudaStreamDestroy(f->own_stream);
    for (auto &hs : f->host_streams) if (hs) cudaStreamDestroy(hs);
}

// src/Filters.jl:284-298, written phase-major: bank[phi*T + (T-1-r)] = h[r*Nphi + phi]
static void make_bank(const std::vector<double> &h, int64_t Nphi, int64_t T, std::vector<double> &bank) {
    bank.assign((size_t)(Nphi * T), 0.0);
    int64_t hIdx = 0;
    const int64_t hLen = (int64_t)h.size();
    for (int64_t row = T - 1; row >= 0; --row)
        for (int64_t col = 0; col < Nphi; ++col, ++hIdx) bank[col * T + row] = hIdx < hLen ? h[hIdx] : 0.0;
}

template <typename R>
static cudaError_t upload_real(const std::vector<double> &src, void **dst) {
    std::vector<R> tmp(src.size());
    for (size_t i = 0; i < src.size(); ++i) tmp[i] = (R)src[i];
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(tmp.size(), 1) * sizeof(R));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, tmp.data(), tmp.size() * sizeof(R), cudaMemcpyHostToDevice);
}

// polyfit(y, order), src/support.jl:85-88: least squares on the Vandermonde matrix A[x, p] = x^p, x = 1..n, p = 0..order,
// coefficients lowest order first.  THE AGREED RECIPE of this library (include/mrb.h, mrb_pfb2pnfb): Householder QR of A
// in Float64, no column scaling or pivoting, back substitution -- the textbook meaning of Julia's `A \ y` for a full-rank
// rectangular A.  cond(A) is ~2.4e6 at order 4 and ~1e8 at order 5 (Nphi = 32): different solvers agree to ~1e-10
// relative, not to the last bit, which is why every binding takes the coefficients from HERE.
static void polyfit_qr(const double *y, int n, int order, double *coef) {
    const int m = order + 1;
    std::vector<double> A((size_t)n * m), b(y, y + n);
    for (int i = 0; i < n; ++i) {
        double p = 1.0;
        for (int j = 0; j < m; ++j) { A[(size_t)i * m + j] = p; p *= (double)(i + 1); }
    }
    for (int k = 0; k < m && k < n; ++k) {
        double norm = 0.0;
        for (int i = k; i < n; ++i) norm += A[(size_t)i * m + k] * A[(size_t)i * m + k];
        norm = std::sqrt(norm);
        if (norm == 0.0) continue;
        const double alpha = A[(size_t)k * m + k] > 0.0 ? -norm : norm;
        std::vector<double> v(n - k);
        for (int i = k; i < n; ++i) v[i - k] = A[(size_t)i * m + k];
        v[0] -= alpha;
        double vv = 0.0;
        for (double e : v) vv += e * e;
        if (vv == 0.0) continue;
        for (int j = k; j < m; ++j) {
            double dot = 0.0;
            for (int i = k; i < n; ++i) dot += v[i - k] * A[(size_t)i * m + j];
            const double s2 = 2.0 * dot / vv;
            for (int i = k; i < n; ++i) A[(size_t)i * m + j] -= s2 * v[i - k];
        }
        double dot = 0.0;
        for (int i = k; i < n; ++i) dot += v[i - k] * b[i];
        const double s2 = 2.0 * dot / vv;
        for (int i = k; i < n; ++i) b[i] -= s2 * v[i - k];
    }
    for (int k = m - 1; k >= 0; --k) {
        double acc = k < n ? b[k] : 0.0;
        for (int j = k + 1; j < m; ++j) acc -= A[(size_t)k * m + j] * coef[j];
        const double d = k < n ? A[(size_t)k * m + k] : 0.0;
        coef[k] = d != 0.0 ? acc / d : 0.0;
    }
}

// pfb2pnfb(pfb, order), src/Filters.jl:311-321: one polynomial per tap ROW of the bank, fitted over phi = 1..Nphi and
// stored as Poly{T}, i.e. rounded to the tap type.  bank is phase-major [Nphi][T]; out is [T][order+1].
static void fit_pnfb(const std::vector<double> &bank, int64_t Nphi, int64_t T, int order, bool tap_f32, double *out) {
    std::vector<double> row((size_t)Nphi);
    for (int64_t i = 0; i < T; ++i) {
        for (int64_t c = 0; c < Nphi; ++c) row[(size_t)c] = bank[(size_t)(c * T + i)];
        double *co = out + i * (order + 1);
        polyfit_qr(row.data(), (int)Nphi, order, co);
        if (tap_f32) for (int p = 0; p <= order; ++p) co[p] = (double)(float)co[p];
    }
}

static void init_state(mrb_filter *f) {
    f->phiIdx = 1; f->deficit = 1; f->xIdx = 1; f->acc = 1.0; f->alpha = 0.0;
}

extern "C" int32_t mrb_create(const mrb_desc *d, mrb_filter **out) {
    if (!d || !out) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    *out = nullptr;
    if (!d->h || d->h_len < 1) return fail(MRB_ERR_BAD_ARGUMENT, "h must hold at least one tap");
    if (d->tap_dtype != MRB_F32 && d->tap_dtype != MRB_F64)
        return fail(MRB_ERR_BAD_ARGUMENT, "tap dtype must be Float32 or Float64");
    if (d->sample_dtype < MRB_F32 || d->sample_dtype > MRB_C128) return fail(MRB_ERR_BAD_ARGUMENT, "bad sample dtype");
    if (d->n_channels < 1) return fail(MRB_ERR_BAD_ARGUMENT, "n_channels must be >= 1");

    std::unique_ptr<mrb_filter> f(new mrb_filter());
    f->th = d->tap_dtype; f->tx = d->sample_dtype; f->ty = promote(f->th, f->tx);
    f->device = d->device; f->nch = d->n_channels; f->hLen = d->h_len;
    f->rate = 0.0; f->delta = 0.0; f->polyorder = -1; f->L = 1; f->M = 1;

    std::vector<double> h((size_t)d->h_len);
    for (int64_t i = 0; i < d->h_len; ++i)
        h[i] = f->th == MRB_F32 ? (double)static_cast<const float *>(d->h)[i] : static_cast<const double *>(d->h)[i];

    int kind = d->kind;
    if (d->rate != 0.0 || kind == MRB_ARBITRARY || kind == MRB_FARROW) {
        // FIRFilter(h, rate, Nphi[, polyorder]) -- src/Filters.jl:183-198
        if (!(d->rate > 0.0)) return fail(MRB_ERR_BAD_ARGUMENT, "rate must be greater than 0");
        const int want = d->poly_order >= 0 ? MRB_FARROW : MRB_ARBITRARY;
        if (kind == MRB_KIND_AUTO) kind = want;
        if (kind != want) return fail(MRB_ERR_BAD_ARGUMENT, "kind does not agree with rate / poly_order");
        f->Nphi = d->n_phi > 0 ? d->n_phi : 32;
        f->rate = d->rate;
        f->delta = (double)f->Nphi / d->rate;                          // :113,142
        f->T = ceil_div(d->h_len, f->Nphi);
        make_bank(h, f->Nphi, f->T, f->bank);
        if (kind == MRB_ARBITRARY) {
            std::vector<double> dh(h.size(), 0.0);                     // dh = [diff(h); 0] in Th arithmetic, :106
            for (size_t i = 0; i + 1 < h.size(); ++i)
                dh[i] = f->th == MRB_F32 ? (double)((float)h[i + 1] - (float)h[i]) : h[i + 1] - h[i];
            make_bank(dh, f->Nphi, f->T, f->dbank);
        } else {
            f->polyorder = d->poly_order;
            if (d->poly_coeffs) {
                f->pnfb.assign(d->poly_coeffs, d->poly_coeffs + f->T * (d->poly_order + 1));
            } else {                                                   // the library's own fit (mrb_pfb2pnfb)
                f->pnfb.assign((size_t)(f->T * (d->poly_order + 1)), 0.0);
                fit_pnfb(f->bank, f->Nphi, f->T, d->poly_order, f->th == MRB_F32, f->pnfb.data());
            }
        }
    } else {
        // FIRFilter(h, ratio) -- src/Filters.jl:158-180
        int64_t L = d->interpolation, M = d->decimation;
        if (L < 1 || M < 1) return fail(MRB_ERR_BAD_ARGUMENT, "interpolation and decimation must be >= 1");
        int64_t a = L, b = M;
        while (b) { int64_t t = a % b; a = b; b = t; }
        L /= a; M /= a;                                                 // Julia Rational is always reduced
        const int want = (L == 1 && M == 1) ? MRB_STANDARD : L == 1 ? MRB_DECIMATOR : M == 1 ? MRB_INTERPOLATOR : MRB_RATIONAL;
        if (kind == MRB_KIND_AUTO) kind = want;
        if (kind != want) return fail(MRB_ERR_BAD_ARGUMENT, "kind does not agree with the resampling ratio");
        f->L = L; f->M = M;
        if (kind == MRB_STANDARD || kind == MRB_DECIMATOR) {
            f->Nphi = 1; f->T = d->h_len;                               // flipud(h), :21,:53
            f->bank.resize(h.size());
            for (size_t i = 0; i < h.size(); ++i) f->bank[i] = h[h.size() - 1 - i];
        } else {
            f->Nphi = L; f->T = ceil_div(d->h_len, L);                  // taps2pfb(h, L), :36,:73
            make_bank(h, f->Nphi, f->T, f->bank);
        }
    }
    f->kind = kind;
    f->H = f->T - 1;                                                    // :165,168,171,174,186,195
    init_state(f.get());

    if (f->device >= 0) {
        int ndev = 0;
        CU(cudaGetDeviceCount(&ndev));
        if (f->device >= ndev) return fail(MRB_ERR_NO_DEVICE, "device %d not present (%d devices)", f->device, ndev);
        DeviceGuard guard(f->device);
        // cudaGetDeviceProperties costs milliseconds: the one-shot filt(h, x, ratio) creates a handle per call, so the
        // properties are looked up once per device and process
        static std::mutex prop_mu;
        static std::vector<std::pair<int, cudaDeviceProp>> prop_cache;
        cudaDeviceProp prop;
        {
            std::lock_guard<std::mutex> lk(prop_mu);
            bool hit = false;
            for (auto &e : prop_cache) if (e.first == f->device) { prop = e.second; hit = true; break; }
            if (!hit) {
                CU(cudaGetDeviceProperties(&prop, f->device));
                prop_cache.emplace_back(f->device, prop);
            }
        }
        if (prop.major != 10)
            return fail(MRB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library holds sm_100a code only", f->device,
                        prop.major, prop.minor);
        f->num_sms = prop.multiProcessorCount;
        const bool dbl = is_double(f->ty);
        CU(dbl ? upload_real<double>(f->bank, &f->d_bank) : upload_real<float>(f->bank, &f->d_bank));
        if (kind == MRB_ARBITRARY) CU(dbl ? upload_real<double>(f->dbank, &f->d_dbank) : upload_real<float>(f->dbank, &f->d_dbank));
        if (kind == MRB_FARROW) {
            CU(cudaMalloc(&f->d_pnfb, f->pnfb.size() * sizeof(double)));
            CU(cudaMemcpy(f->d_pnfb, f->pnfb.data(), f->pnfb.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        const size_t hb = std::max<size_t>((size_t)(f->H * f->nch) * dsize(f->tx), 16);
        for (int i = 0; i < 2; ++i) { CU(cudaMalloc(&f->d_hist[i], hb)); CU(cudaMemset(f->d_hist[i], 0, hb)); }
        CU(cudaStreamCreateWithFlags(&f->own_stream, cudaStreamNonBlocking));
        int32_t rc = tiled_prepare(f->tiled, kind, f->tx, f->ty, f->L, f->M, f->Nphi, f->T, f->bank, f->dbank, prop);
        if (rc != 0) return fail(MRB_ERR_CUDA, "tiled_prepare failed: %s", cudaGetErrorString((cudaError_t)rc));
        rc = unit_prepare(f->unit, kind, f->tx, f->ty, f->L, f->M, f->T, f->bank, prop);
        if (rc != 0) return fail(MRB_ERR_CUDA, "unit_prepare failed: %s", cudaGetErrorString((cudaError_t)rc));
        rc = decim_prepare(f->decim, kind, f->tx, f->ty, f->L, f->M, f->T, f->bank, prop);
        if (rc != 0) return fail(MRB_ERR_CUDA, "decim_prepare failed: %s", cudaGetErrorString((cudaError_t)rc));
        rc = table_prepare(f->table, kind, f->tx, f->ty, f->T, (kind == MRB_ARBITRARY || kind == MRB_FARROW) ? f->rate : (double)f->L / (double)f->M, prop);
        if (rc != 0) return fail(MRB_ERR_CUDA, "table_prepare failed: %s", cudaGetErrorString((cudaError_t)rc));
        rc = mma_prepare(f->mma, kind, f->tx, f->ty, f->th, f->T, prop);
        if (rc != 0) return fail(MRB_ERR_CUDA, "mma_prepare failed: %s", cudaGetErrorString((cudaError_t)rc));
        CU(cudaDeviceSynchronize());
    }
    *out = f.release();
    return MRB_OK;
}

// Bank construction ON THE DEVICE (src/Filters.jl:21,53 flipud, :284-298 taps2pfb, :106 [diff(h); 0]):
// from the raw taps h (tap dtype TH) to the phase-major bank rows (compute type R) the kernels read from global memory.
//   bank[phi*T + (T-1-r)] = h[r*Nphi + phi]   (zero past hLen);   Nphi = 1: bank[i] = h[hLen-1-i]
//   dbank: the same of dh[i] = h[i+1] - h[i] in TH arithmetic, dh[hLen-1] = 0
template <typename TH, typename R>
__global__ void __launch_bounds__(256) k_build_banks(const TH *__restrict__ h, int64_t hLen, int64_t Nphi, int64_t T, R *__restrict__ bank,
                                                     R *__restrict__ dbank, float *__restrict__ bank_f32) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= Nphi * T) return;
    const int64_t phi = idx / T, row = idx - phi * T;
    const int64_t hi = (T - 1 - row) * Nphi + phi;                   // make_bank: rows are filled from the last one up
    const TH v = hi < hLen ? h[hi] : TH(0);
    bank[idx] = (R)v;
    if (bank_f32) bank_f32[idx] = (float)v;
    if (dbank) {
        TH d = TH(0);
        if (hi + 1 < hLen) d = h[hi + 1] - h[hi];                    // [diff(h); 0] in the tap type, :106
        else if (hi < hLen) d = TH(0);
        dbank[idx] = (R)d;
    }
}

// Live tap update: replace the taps of a filter in place -- same length and tap dtype, so T, H, the
// carried phase state and the per-channel history all stay as they are -- e.g. an adaptive filter whose taps change
// between chunks.  ASYNCHRONOUS and ordered on `stream`: no device-wide synchronisation.
//  * kernels that keep their taps in the kernel PARAMETER block (tiled, unit) are launched with a copy of the host block,
//    so rewriting the host block is all they need;
//  * the banks in global memory (d_bank, d_dbank: generic / stream / head / table / tensor-core kernels; the decimator's
//    residue tables; the Farrow coefficients) are rebuilt ON THE DEVICE by k_build_banks / k_decim_taps from the raw
//    taps, which travel through a small pinned staging buffer; work already queued on the stream still sees the old banks.
// Host banks are built into temporaries and committed only when every step succeeded.
static int32_t set_taps_impl(mrb_filter *f, const void *hv, int64_t h_len, const double *poly_coeffs, cudaStream_t st) {
    if (!f || !hv) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    if (h_len != f->hLen) return fail(MRB_ERR_BAD_ARGUMENT, "mrb_set_taps keeps the tap count (%lld), got %lld", (long long)f->hLen, (long long)h_len);
    std::vector<double> h((size_t)h_len), bank, dbank, pnfb;
    for (int64_t i = 0; i < h_len; ++i)
        h[i] = f->th == MRB_F32 ? (double)static_cast<const float *>(hv)[i] : static_cast<const double *>(hv)[i];
    if (f->kind == MRB_STANDARD || f->kind == MRB_DECIMATOR) {
        bank.resize(h.size());
        for (size_t i = 0; i < h.size(); ++i) bank[i] = h[h.size() - 1 - i];               // flipud(h), :21,:53
    } else {
        make_bank(h, f->Nphi, f->T, bank);                                                 // taps2pfb, :36,:73,:107,:138
    }
    if (f->kind == MRB_ARBITRARY) {
        std::vector<double> dh(h.size(), 0.0);                                             // :106
        for (size_t i = 0; i + 1 < h.size(); ++i)
            dh[i] = f->th == MRB_F32 ? (double)((float)h[i + 1] - (float)h[i]) : h[i + 1] - h[i];
        make_bank(dh, f->Nphi, f->T, dbank);
    }
    if (f->kind == MRB_FARROW) {
        pnfb.resize(f->pnfb.size());
        if (poly_coeffs) pnfb.assign(poly_coeffs, poly_coeffs + f->T * (f->polyorder + 1));
        else fit_pnfb(bank, f->Nphi, f->T, f->polyorder, f->th == MRB_F32, pnfb.data());
    }
    if (f->device >= 0) {
        DeviceGuard guard(f->device);
        int32_t rc = order_after_last(f, st);
        if (rc) return rc;
        const size_t hb = (size_t)h_len * (f->th == MRB_F32 ? 4 : 8), pb = pnfb.size() * sizeof(double);
        if (f->taps_stage_bytes < hb + pb) {
            if (f->taps_pending) { CU(cudaEventSynchronize(f->taps_ev)); f->taps_pending = false; }
            cudaFreeHost(f->h_taps_stage); cudaFree(f->d_taps_stage);
            f->h_taps_stage = nullptr; f->d_taps_stage = nullptr; f->taps_stage_bytes = 0;
            CU(cudaMallocHost(&f->h_taps_stage, hb + pb + 64));
            CU(cudaMalloc(&f->d_taps_stage, hb + 64));
            f->taps_stage_bytes = hb + pb;
            if (!f->taps_ev) CU(cudaEventCreateWithFlags(&f->taps_ev, cudaEventDisableTiming));
        }
        if (f->taps_pending) { CU(cudaEventSynchronize(f->taps_ev)); f->taps_pending = false; }   // the last update's copies have left the staging buffer
        memcpy(f->h_taps_stage, hv, hb);
        CU(cudaMemcpyAsync(f->d_taps_stage, f->h_taps_stage, hb, cudaMemcpyHostToDevice, st));
        if (pb) {
            memcpy(static_cast<char *>(f->h_taps_stage) + hb, pnfb.data(), pb);
            CU(cudaMemcpyAsync(f->d_pnfb, static_cast<char *>(f->h_taps_stage) + hb, pb, cudaMemcpyHostToDevice, st));
        }
        CU(cudaEventRecord(f->taps_ev, st));
        f->taps_pending = true;
        const bool dbl = is_double(f->ty);
        const int64_t nb = f->Nphi * f->T;
        const unsigned g = (unsigned)ceil_div(nb, 256);
        float *bank_f32 = nullptr;                                     // the decimator tables are built from float32 rows
        if (f->decim.ok && !dbl) bank_f32 = static_cast<float *>(f->d_bank);
        if (f->th == MRB_F32) {
            if (dbl) k_build_banks<float, double><<<g, 256, 0, st>>>((const float *)f->d_taps_stage, h_len, f->Nphi, f->T, (double *)f->d_bank, (double *)f->d_dbank, nullptr);
            else k_build_banks<float, float><<<g, 256, 0, st>>>((const float *)f->d_taps_stage, h_len, f->Nphi, f->T, (float *)f->d_bank, (float *)f->d_dbank, nullptr);
        } else {
            if (dbl) k_build_banks<double, double><<<g, 256, 0, st>>>((const double *)f->d_taps_stage, h_len, f->Nphi, f->T, (double *)f->d_bank, (double *)f->d_dbank, nullptr);
            else k_build_banks<double, float><<<g, 256, 0, st>>>((const double *)f->d_taps_stage, h_len, f->Nphi, f->T, (float *)f->d_bank, (float *)f->d_dbank, nullptr);
        }
        ++f->launches;
        if (bank_f32) { decim_set_taps(f->decim, bank_f32, st); ++f->launches; }
        CU(cudaGetLastError());
        // parameter-block kernels: their host blocks (copied into every launch)
        tiled_set_bank(f->tiled, f->L, f->T, bank);
        unit_set_bank(f->unit, f->T, bank);
        decim8_set_bank(f->decim, bank);
        f->last_stream = st; f->last_valid = true;
    }
    f->bank.swap(bank);
    if (f->kind == MRB_ARBITRARY) f->dbank.swap(dbank);
    if (f->kind == MRB_FARROW) f->pnfb.swap(pnfb);
    return MRB_OK;
}

extern "C" int32_t mrb_set_taps_async(mrb_filter *f, const void *hv, int64_t h_len, const double *poly_coeffs, void *stream) {
    return set_taps_impl(f, hv, h_len, poly_coeffs, (cudaStream_t)stream);
}

// The same on the stream of the handle's last asynchronous call (the default stream before any)
extern "C" int32_t mrb_set_taps(mrb_filter *f, const void *hv, int64_t h_len, const double *poly_coeffs) {
    return set_taps_impl(f, hv, h_len, poly_coeffs, f && f->last_valid ? f->last_stream : (cudaStream_t) nullptr);
}

extern "C" int32_t mrb_destroy(mrb_filter *f) {
    delete f;                                              // ~mrb_filter releases the device side
    return MRB_OK;
}

extern "C" int32_t mrb_get_info(const mrb_filter *f, mrb_info *o) {
    if (!f || !o) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    o->kind = f->kind; o->tap_dtype = f->th; o->sample_dtype = f->tx; o->out_dtype = f->ty; o->device = f->device;
    o->n_phi = (int32_t)f->Nphi; o->poly_order = f->polyorder;
    o->taps_per_phase = f->T; o->history_len = f->H; o->h_len = f->hLen;
    o->interpolation = f->L; o->decimation = f->M; o->n_channels = f->nch; o->rate = f->rate;
    return MRB_OK;
}

// ------------------------------------------------------------------------------------------
// sequencing
// ------------------------------------------------------------------------------------------
static bool is_table_kind(const mrb_filter *f) { return f->kind == MRB_ARBITRARY || f->kind == MRB_FARROW; }

extern "C" int32_t mrb_outputlength(const mrb_filter *f, int64_t n_in, int64_t *n_out) {
    if (!f || !n_out || n_in < 0) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    switch (f->kind) {
    case MRB_STANDARD: *n_out = n_in; break;                                              // :359-361
    case MRB_INTERPOLATOR: *n_out = f->L * n_in; break;                                   // :363-365
    case MRB_DECIMATOR:                                                                   // :367-369
    case MRB_RATIONAL:                                                                    // :371-373 -> :352-357
        *n_out = (int64_t)std::ceil((double)((n_in - f->deficit + 1) * f->L - f->phiIdx + 1) / (double)f->M);
        break;
    default: *n_out = (int64_t)std::ceil((double)(n_in - f->deficit + 1) * f->rate); break;   // :375-381
    }
    return MRB_OK;
}

// Exact replay for the table kinds.  Fills the store (n 0-based, phi 0-based, alpha | Float64 phase) when one is given.
// Returns the output count, or -1 when the store could not grow.
static int64_t replay_table(const mrb_filter *f, int64_t n_in, mrb_state *end, SchedStore *st) {
    mrb_state s{f->phiIdx, f->deficit, f->xIdx, f->acc, f->alpha};
    int64_t count = 0;
    if (st) st->size = 0;
    if (n_in < s.input_deficit) {                                       // :705-709, :805-809
        s.input_deficit -= n_in;
    } else {
        ArbState a{s.phi_accumulator, s.input_deficit};                 // xIdx = inputDeficit, :715,:812
        const bool arb = f->kind == MRB_ARBITRARY;
        const ArbStepper stepper(f->delta, f->Nphi);
        if (st) {
            // schedule wanted: written through raw pointers into storage sized for the bound of outputlength (:375-381)
            // plus slack, grown if the loop decides otherwise
            const bool pin = f->device >= 0;
            size_t cap = (size_t)((double)(n_in - s.input_deficit + 1) * f->rate) + 16;
            if (!st->reserve(cap, pin)) return -1;
            cap = st->cap;
            int64_t *pn = st->n; double *pa = st->a; int32_t *pp = st->phi;
            size_t c = 0;
            const int64_t phi0 = s.phi_idx;
            const double alpha0 = s.alpha;
            while (a.xIdx <= n_in) {
                if (c == cap) {
                    st->size = c;
                    if (!st->reserve(cap + cap / 2 + 16, pin)) return -1;
                    cap = st->cap; pn = st->n; pa = st->a; pp = st->phi;
                }
                pn[c] = a.xIdx - 1;
                pa[c] = a.acc;               // farrow: the Float64 phiIdx the taps are evaluated at; arbitrary: split below
                ++c;
                stepper.step(a);
            }
            if (arb) {
                // branch and alpha of every output from the accumulator it started with (:671-672) -- outside the
                // sequential loop, where the conversion vectorises.  Output 0 keeps the carried pair: setphase may
                // have clamped it.
                if (c > 0) { pp[0] = (int32_t)(phi0 - 1); pa[0] = alpha0; }
                for (size_t i = 1; i < c; ++i) {
                    const int32_t ph = (int32_t)pa[i];                  // floor of a value in [1, Nphi+1)
                    pp[i] = ph - 1;
                    pa[i] -= (double)ph;
                }
                s.phi_idx = (int64_t)a.acc;
                s.alpha = a.acc - (double)s.phi_idx;
            }
            st->size = c;
            count = (int64_t)c;
        } else {
            while (a.xIdx <= n_in) {
                ++count;
                stepper.step(a);
            }
            if (arb && count > 0) {
                s.phi_idx = (int64_t)a.acc;
                s.alpha = a.acc - (double)s.phi_idx;
            }
        }
        s.phi_accumulator = a.acc;
        s.x_idx = a.xIdx;
        s.input_deficit = a.xIdx - n_in;                                // :734,:828
    }
    if (end) *end = s;
    return count;
}

// table kinds: replay (or reuse the cached replay of) the schedule of n_in inputs from the current state
static int64_t replay_cached(mrb_filter *f, int64_t n_in, mrb_state *end) {
    const mrb_state from{f->phiIdx, f->deficit, f->xIdx, f->acc, f->alpha};
    if (!(f->sched_valid && f->sched_n_in == n_in && memcmp(&from, &f->sched_from, sizeof from) == 0)) {
        f->sched_cur ^= 1;                                          // the other store: uploads of the last call may still read this one
        SchedStore &st = f->sched[f->sched_cur];
        if (st.pending) {
            DeviceGuard guard(f->device);
            cudaEventSynchronize(st.ev);
            st.pending = false;
        }
        f->sched_valid = false;
        f->sched_count = replay_table(f, n_in, &f->sched_end, &st);
        if (f->sched_count < 0) { f->sched_count = 0; st.size = 0; return -1; }
        f->sched_from = from; f->sched_n_in = n_in; f->sched_valid = true;
    }
    if (end) *end = f->sched_end;
    return f->sched_count;
}

static int64_t count_outputs(const mrb_filter *f, int64_t n_in, mrb_state *end) {
    if (is_table_kind(f)) return replay_cached(const_cast<mrb_filter *>(f), n_in, end);
    int64_t p = f->phiIdx - 1, dd = f->deficit;
    const int64_t N = IntSeq::count(f->L, f->M, p, dd, n_in);
    IntSeq::advance(f->L, f->M, p, dd, n_in);
    if (end) { *end = mrb_state{p + 1, dd, f->xIdx, f->acc, f->alpha}; }
    return N;
}

extern "C" int32_t mrb_output_count(const mrb_filter *f, int64_t n_in, int64_t *n_out) {
    if (!f || !n_out || n_in < 0) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    *n_out = count_outputs(f, n_in, nullptr);
    if (*n_out < 0) return fail(MRB_ERR_CUDA, "out of memory for the schedule");
    return MRB_OK;
}

static void commit_state(mrb_filter *f, const mrb_state &s) {
    f->phiIdx = s.phi_idx; f->deficit = s.input_deficit; f->xIdx = s.x_idx; f->acc = s.phi_accumulator; f->alpha = s.alpha;
}

// The schedule the next n_in inputs will produce, without advancing the state: per output the 0-based index of the
// window's last input sample, and for the arbitrary-rate kinds the 0-based branch (arbitrary) and alpha (arbitrary) or
// Float64 phase (farrow).  Any of the three destinations may be NULL; each must hold mrb_output_count(n_in) entries.
extern "C" int32_t mrb_get_schedule(mrb_filter *f, int64_t n_in, int64_t *n_idx, int32_t *branch, double *frac) {
    if (!f || n_in < 0) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    if (is_table_kind(f)) {
        const int64_t N = replay_cached(f, n_in, nullptr);
        if (N < 0) return fail(MRB_ERR_CUDA, "out of memory for the schedule");
        if (N == 0) return MRB_OK;
        const SchedStore &st = f->sched[f->sched_cur];
        if (n_idx) memcpy(n_idx, st.n, (size_t)N * sizeof(int64_t));
        if (frac) memcpy(frac, st.a, (size_t)N * sizeof(double));
        if (branch) {
            if (f->kind == MRB_ARBITRARY) memcpy(branch, st.phi, (size_t)N * sizeof(int32_t));
            else for (int64_t k = 0; k < N; ++k) branch[k] = (int32_t)st.a[k] - 1;
        }
        return MRB_OK;
    }
    const int64_t p = f->phiIdx - 1, d = f->deficit;
    const int64_t N = IntSeq::count(f->L, f->M, p, d, n_in);
    for (int64_t k = 0; k < N; ++k) {                                    // closed form of :558-569, :613-625
        const int64_t t = p + k * f->M;
        if (n_idx) n_idx[k] = d - 1 + t / f->L;
        if (branch) branch[k] = (int32_t)(t % f->L);
        if (frac) frac[k] = 0.0;
    }
    return MRB_OK;
}

extern "C" int32_t mrb_advance(mrb_filter *f, int64_t n_in, int64_t *n_out) {
    if (!f || n_in < 0) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    mrb_state e;
    const int64_t N = count_outputs(f, n_in, &e);
    if (N < 0) return fail(MRB_ERR_CUDA, "out of memory for the schedule");
    commit_state(f, e);
    if (n_out) *n_out = N;
    return MRB_OK;
}

extern "C" int32_t mrb_inputlength(int64_t n_out, int64_t L, int64_t M, int64_t phi, int64_t *n_in) {
    if (!n_in || L < 1 || M < 1) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    *n_in = (int64_t)std::ceil((double)(n_out * M + phi - 1) / (double)L);                // :396-401
    return MRB_OK;
}

extern "C" int32_t mrb_nextphase(int64_t cur, int64_t L, int64_t M, int64_t *next) {
    if (!next || L < 1 || M < 1) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    const int64_t nx = cur + M % L;                                                       // :436-438
    *next = nx > L ? nx - L : nx;
    return MRB_OK;
}

extern "C" int32_t mrb_pfb2pnfb(const void *h, int64_t h_len, int32_t dtype, int64_t n_phi, int32_t order, double *coeffs) {
    if (!h || !coeffs || h_len < 1 || n_phi < 1 || order < 0 || (dtype != MRB_F32 && dtype != MRB_F64))
        return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    std::vector<double> hv((size_t)h_len), bank;
    for (int64_t i = 0; i < h_len; ++i)
        hv[(size_t)i] = dtype == MRB_F32 ? (double)static_cast<const float *>(h)[i] : static_cast<const double *>(h)[i];
    const int64_t T = ceil_div(h_len, n_phi);
    make_bank(hv, n_phi, T, bank);
    fit_pnfb(bank, n_phi, T, order, dtype == MRB_F32, coeffs);
    return MRB_OK;
}

extern "C" int32_t mrb_taps2pfb(const void *h, int64_t h_len, int32_t dtype, int64_t n_phi, void *pfb) {
    if (!h || !pfb || h_len < 1 || n_phi < 1 || (dtype != MRB_F32 && dtype != MRB_F64))
        return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    const int64_t T = ceil_div(h_len, n_phi);
    int64_t hIdx = 0;
    for (int64_t row = T - 1; row >= 0; --row)
        for (int64_t col = 0; col < n_phi; ++col, ++hIdx) {
            if (dtype == MRB_F32) static_cast<float *>(pfb)[col * T + row] = hIdx < h_len ? static_cast<const float *>(h)[hIdx] : 0.f;
            else static_cast<double *>(pfb)[col * T + row] = hIdx < h_len ? static_cast<const double *>(h)[hIdx] : 0.0;
        }
    return MRB_OK;
}

// ------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------
extern "C" int32_t mrb_reset(mrb_filter *f) {
    if (!f) return fail(MRB_ERR_BAD_ARGUMENT, "null handle");
    init_state(f);
    if (f->device >= 0) {
        DeviceGuard guard(f->device);
        const size_t hb = (size_t)(f->H * f->nch) * dsize(f->tx);
        CU(cudaDeviceSynchronize());
        if (hb) CU(cudaMemset(f->d_hist[f->cur], 0, hb));
        CU(cudaDeviceSynchronize());
    }
    return MRB_OK;
}

extern "C" int32_t mrb_setphase(mrb_filter *f, double phi) {
    if (!f) return fail(MRB_ERR_BAD_ARGUMENT, "null handle");
    if (!(phi >= 0.0 && phi <= 1.0)) return fail(MRB_ERR_BAD_ARGUMENT, "phase must be >= 0 and <= 1");   // :211,217,225
    switch (f->kind) {
    case MRB_RATIONAL:
        f->phiIdx = std::min<int64_t>((int64_t)std::floor(phi * (double)f->Nphi) + 1, f->Nphi);
        return MRB_OK;
    case MRB_ARBITRARY: {
        f->acc = 1.0 + phi * (double)f->Nphi;
        f->phiIdx = std::min<int64_t>((int64_t)std::floor(f->acc), f->Nphi);
        f->alpha = f->acc - (double)f->phiIdx;
        return MRB_OK;
    }
    case MRB_FARROW:
        f->acc = phi * (double)(f->Nphi - 1) + 1.0;                      // :226
        return MRB_OK;
    default: return fail(MRB_ERR_UNSUPPORTED, "setphase is not supported for this kernel (it carries no phase)");
    }
}

extern "C" int32_t mrb_get_state(const mrb_filter *f, mrb_state *s) {
    if (!f || !s) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    *s = mrb_state{f->phiIdx, f->deficit, f->xIdx, f->acc, f->alpha};
    return MRB_OK;
}

extern "C" int32_t mrb_set_state(mrb_filter *f, const mrb_state *s) {
    if (!f || !s) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    const int64_t maxphi = is_table_kind(f) ? f->Nphi : f->L;
    if (s->phi_idx < 1 || s->phi_idx > maxphi || s->input_deficit < 1)
        return fail(MRB_ERR_BAD_ARGUMENT, "state out of range");
    // closed upper bound: setphase(1.0) leaves the accumulator AT Nphi+1 (the next update wraps it), and a state read
    // back from one handle must be accepted by another
    if (is_table_kind(f) && !(s->phi_accumulator >= 1.0 && s->phi_accumulator <= (double)(f->Nphi + 1)))
        return fail(MRB_ERR_BAD_ARGUMENT, "phase accumulator out of range");
    commit_state(f, *s);
    return MRB_OK;
}

extern "C" int32_t mrb_get_history(mrb_filter *f, void *dst) {
    if (!f || !dst) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    if (f->device < 0) return fail(MRB_ERR_NO_DEVICE, "host-only handle holds no history");
    DeviceGuard guard(f->device);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(dst, f->d_hist[f->cur], (size_t)(f->H * f->nch) * dsize(f->tx), cudaMemcpyDeviceToHost));
    return MRB_OK;
}

extern "C" int32_t mrb_set_history(mrb_filter *f, const void *src) {
    if (!f || !src) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    if (f->device < 0) return fail(MRB_ERR_NO_DEVICE, "host-only handle holds no history");
    DeviceGuard guard(f->device);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(f->d_hist[f->cur], src, (size_t)(f->H * f->nch) * dsize(f->tx), cudaMemcpyHostToDevice));
    return MRB_OK;
}

extern "C" int32_t mrb_tapsforphase(const mrb_filter *f, double phase, void *taps) {
    if (!f || !taps) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    if (!is_table_kind(f)) return fail(MRB_ERR_UNSUPPORTED, "tapsforphase is defined for arbitrary and farrow kernels");
    if (!(phase >= 0.0 && phase <= (double)(f->Nphi + 1)))
        return fail(MRB_ERR_BAD_ARGUMENT, "phase must be >= 0 and <= Nphi+1");            // :678,:765
    for (int64_t i = 0; i < f->T; ++i) {
        double v;
        if (f->kind == MRB_ARBITRARY) {                                                   // :681-686
            double ip;
            const double a = std::modf(phase, &ip);
            const int64_t phiIdx = (int64_t)ip;
            if (phiIdx < 1 || phiIdx > f->Nphi) return fail(MRB_ERR_BAD_ARGUMENT, "phase selects branch %lld outside 1..Nphi", (long long)phiIdx);
            v = f->bank[(phiIdx - 1) * f->T + i] + a * f->dbank[(phiIdx - 1) * f->T + i];
        } else {                                                                          // :768-770
            const double *c = &f->pnfb[i * (f->polyorder + 1)];
            v = c[f->polyorder];
            for (int p = f->polyorder - 1; p >= 0; --p) v = v * phase + c[p];
        }
        if (f->th == MRB_F32) static_cast<float *>(taps)[i] = (float)v;
        else static_cast<double *>(taps)[i] = v;
    }
    return MRB_OK;
}

extern "C" int32_t mrb_get_pfb(const mrb_filter *f, int32_t which, void *dst) {
    if (!f || !dst) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    const std::vector<double> &b = which == 0 ? f->bank : f->dbank;
    if (which < 0 || which > 1 || b.empty()) return fail(MRB_ERR_BAD_ARGUMENT, "no such bank");
    for (size_t i = 0; i < b.size(); ++i) {   // phase-major [phi][T] == Julia column-major T x Nphi
        if (f->th == MRB_F32) static_cast<float *>(dst)[i] = (float)b[i];
        else static_cast<double *>(dst)[i] = b[i];
    }
    return MRB_OK;
}

// ------------------------------------------------------------------------------------------
// filt
// ------------------------------------------------------------------------------------------
template <typename RX, typename R, int NC>
static void launch_generic(const GenParams &P0, cudaStream_t st) {
    GenParams P = P0;
    P.cpb_log2 = 0;
    while (P.cpb_log2 < 5 && (128 >> P.cpb_log2) >= P.nout && (2ll << P.cpb_log2) <= P.nch) ++P.cpb_log2;   // short launch: share the block among channels
    const int64_t W = 256 >> P.cpb_log2, cpb = 1ll << P.cpb_log2;
    dim3 grid((unsigned)ceil_div(P.nout, W), (unsigned)std::min<int64_t>(ceil_div(P.nch, cpb), 32768));
    k_generic<RX, R, NC><<<grid, 256, 0, st>>>(P);
}

// k_stream: integer schedules with enough outputs to pay for staging the bank once per CTA
constexpr int64_t kStreamMinOutputs = 8192;
constexpr int64_t kStreamMaxBankBytes = 96 * 1024;

template <typename RX, typename R, int NC>
static bool launch_stream(const GenParams &P, int num_sms, int device, cudaStream_t st) {
    const int64_t pitch = P.T | 1;                                     // odd row pitch: 32 branches -> 32 banks
    const int64_t bytes = P.L * pitch * (int64_t)sizeof(R);
    if (P.mode != SEQ_INTEGER || P.nout < kStreamMinOutputs || bytes > kStreamMaxBankBytes || P.L >= (1ll << 31) / pitch)
        return false;
    static std::atomic<bool> attr_set[64];                             // per instantiation and device; handles on
    if (device < 0 || device >= 64) return false;                      // different threads may race here: setting the
    if (!attr_set[device].load(std::memory_order_acquire)) {           // attribute twice is harmless, a torn flag is not
        if (cudaFuncSetAttribute(k_stream<RX, R, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamMaxBankBytes) != cudaSuccess)
            return false;
        attr_set[device].store(true, std::memory_order_release);
    }
    const int64_t per_sm = std::max<int64_t>(1, std::min<int64_t>(8, (200 * 1024) / (bytes + 1024)));
    const int64_t target = per_sm * num_sms;
    const int64_t gx = std::min<int64_t>(ceil_div(P.nout, 256), target);
    const int64_t gy = std::min<int64_t>(P.nch, std::max<int64_t>(1, target / gx));
    k_stream<RX, R, NC><<<dim3((unsigned)gx, (unsigned)gy), 256, (size_t)bytes, st>>>(P, (int)pitch);
    return true;
}

// k_head_warp: short launches (chunk heads) of integer schedules with long windows
template <typename RX, typename R, int NC>
static bool launch_head(const GenParams &P, cudaStream_t st) {
    if (P.nout > 256 || P.T < 64 || P.nout * P.nch >= (1ll << 26)) return false;
    // 32x the warps of k_generic: pays when a thread-per-output warp would read M-strided windows (decimating ratios)
    // or when there are too few outputs to fill the machine with threads (measured: standard-128 x 4096 channels loses).
    // Arbitrary-rate heads (two dot products per output, 23 us for 96 outputs x 1024 channels on k_generic) always pay.
    if (P.mode == SEQ_INTEGER && P.M < 4 * P.L && P.nout * P.nch >= 8192) return false;
    k_head_warp<RX, R, NC><<<(unsigned)ceil_div(P.nout * P.nch * 32, 256), 256, 0, st>>>(P);
    return true;
}

// Returns the name of the kernel that was launched.  policy != 0 (MRB_POLICY_GENERIC) keeps to k_generic.
static const char *dispatch_generic(const mrb_filter *f, const GenParams &P, cudaStream_t st) {
    const int key = f->tx * 4 + f->ty;
    const int sms = f->num_sms;
    if (f->policy != 1) {
        bool done = false;
        switch (key) {
        case MRB_F32 * 4 + MRB_F32: done = launch_head<float, float, 1>(P, st); break;
        case MRB_C64 * 4 + MRB_C64: done = launch_head<float, float, 2>(P, st); break;
        case MRB_F32 * 4 + MRB_F64: done = launch_head<float, double, 1>(P, st); break;
        case MRB_C64 * 4 + MRB_C128: done = launch_head<float, double, 2>(P, st); break;
        case MRB_F64 * 4 + MRB_F64: done = launch_head<double, double, 1>(P, st); break;
        case MRB_C128 * 4 + MRB_C128: done = launch_head<double, double, 2>(P, st); break;
        }
        if (done) return "head";
        switch (key) {
        case MRB_F32 * 4 + MRB_F32: done = launch_stream<float, float, 1>(P, sms, f->device, st); break;
        case MRB_C64 * 4 + MRB_C64: done = launch_stream<float, float, 2>(P, sms, f->device, st); break;
        case MRB_F32 * 4 + MRB_F64: done = launch_stream<float, double, 1>(P, sms, f->device, st); break;
        case MRB_C64 * 4 + MRB_C128: done = launch_stream<float, double, 2>(P, sms, f->device, st); break;
        case MRB_F64 * 4 + MRB_F64: done = launch_stream<double, double, 1>(P, sms, f->device, st); break;
        case MRB_C128 * 4 + MRB_C128: done = launch_stream<double, double, 2>(P, sms, f->device, st); break;
        }
        if (done) return "stream";
    }
    switch (key) {
    case MRB_F32 * 4 + MRB_F32: launch_generic<float, float, 1>(P, st); break;
    case MRB_C64 * 4 + MRB_C64: launch_generic<float, float, 2>(P, st); break;
    case MRB_F32 * 4 + MRB_F64: launch_generic<float, double, 1>(P, st); break;
    case MRB_C64 * 4 + MRB_C128: launch_generic<float, double, 2>(P, st); break;
    case MRB_F64 * 4 + MRB_F64: launch_generic<double, double, 1>(P, st); break;
    case MRB_C128 * 4 + MRB_C128: launch_generic<double, double, 2>(P, st); break;
    }
    return "generic";
}

static void launch_history(const mrb_filter *f, const void *x, int64_t ldx, int64_t n_in, const void *hold, void *hnew,
                           int64_t nch, cudaStream_t st) {
    if (f->H == 0) return;
    const unsigned g = (unsigned)ceil_div(f->H * nch, 256);
    switch (f->tx) {
    case MRB_F32: k_history<float, 1><<<g, 256, 0, st>>>((const float *)x, ldx, n_in, (const float *)hold, (float *)hnew, f->H, nch); break;
    case MRB_F64: k_history<double, 1><<<g, 256, 0, st>>>((const double *)x, ldx, n_in, (const double *)hold, (double *)hnew, f->H, nch); break;
    case MRB_C64: k_history<float, 2><<<g, 256, 0, st>>>((const float *)x, ldx, n_in, (const float *)hold, (float *)hnew, f->H, nch); break;
    case MRB_C128: k_history<double, 2><<<g, 256, 0, st>>>((const double *)x, ldx, n_in, (const double *)hold, (double *)hnew, f->H, nch); break;
    }
}

// side stream of a table context: the chunk head runs there beside the main kernel (MRB_NO_SIDE_STREAM=1: same stream)
static int32_t ensure_side(TableCtx &c) {
    static const bool no_side = getenv("MRB_NO_SIDE_STREAM") != nullptr;
    if (no_side || c.side) return MRB_OK;
    CU(cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    return MRB_OK;
}

static int32_t ensure_sched(mrb_filter *f, TableCtx &c) {
    if (c.ready) return MRB_OK;
    for (auto &s : c.slot) {
        CU(cudaMalloc(&s.d_n, kSchedChunk * sizeof(int64_t)));
        CU(cudaMalloc(&s.d_phi, kSchedChunk * sizeof(int32_t)));
        CU(cudaMalloc(&s.d_a, kSchedChunk * sizeof(double)));
    }
    if (f->kind == MRB_FARROW)
        CU(cudaMalloc(&c.d_taptab, (size_t)kSchedChunk * f->T * (is_double(f->ty) ? 8 : 4)));
    if (!f->mma.ok) {                  // (the tensor-core kernel computes its head itself)
        int32_t rc = ensure_side(c);
        if (rc) return rc;
    }
    c.ready = true;
    return MRB_OK;
}

// A call on stream `st` must not overtake work the handle still has in flight on another stream (history double buffer,
// schedule slices, tap rows): make `st` wait for it on the device.  No host synchronisation.
static int32_t order_after_last(mrb_filter *f, cudaStream_t st) {
    if (f->last_valid && f->last_stream != st) {
        if (!f->order_ev) CU(cudaEventCreateWithFlags(&f->order_ev, cudaEventDisableTiming));
        CU(cudaEventRecord(f->order_ev, f->last_stream));
        CU(cudaStreamWaitEvent(st, f->order_ev, 0));
    }
    return MRB_OK;
}

// Filter channels [c0, c0+nc) of one chunk from the CURRENT handle state (not committed here).
// x / y point at channel c0.  N = exact output count for this chunk.
// `ctx` selects the table context (schedule slices, tap rows) this call may write: one per concurrently running stream.
static int32_t run_channels(mrb_filter *f, const void *x, int64_t ldx, int64_t n_in, void *y, int64_t ldy, int64_t N,
                            int64_t c0, int64_t nc, cudaStream_t st, int ctx) {
    const size_t es = dsize(f->tx);
    const char *hold = static_cast<const char *>(f->d_hist[f->cur]) + (size_t)(c0 * f->H) * es;
    char *hnew = static_cast<char *>(f->d_hist[f->cur ^ 1]) + (size_t)(c0 * f->H) * es;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (f->timing && N > 0) {
        CU(cudaEventCreate(&t0)); CU(cudaEventCreate(&t1));
        CU(cudaEventRecord(t0, st));
    }
    if (N > 0) {
        GenParams P{};
        P.x = x; P.ldx = ldx; P.n_in = n_in; P.hist = hold; P.H = f->H; P.y = y; P.ldy = ldy;
        P.bank = f->d_bank; P.dbank = f->d_dbank; P.T = f->T; P.nch = nc;
        if (!is_table_kind(f)) {
            P.mode = SEQ_INTEGER; P.L = f->L; P.M = f->M; P.p0 = f->phiIdx - 1; P.d0m1 = f->deficit - 1;
            P.k_base = 0; P.nout = N;
            int64_t k_begin = -1;
            if (f->policy == 0 && f->mma.ok) {
                // tensor cores (float32 samples and taps): a group of 32 outputs spans at most floor(31 M / L) + 1 inputs
                static const bool mma_int = !(getenv("MRB_MMA_INT") && atoi(getenv("MRB_MMA_INT")) == 0);
                // short single-rate filters are HBM bound on the CUDA cores already (standard, 32 taps: k_unit_f32 730 Gout/s,
                // tensor-core kernel 592); from ~48 taps on the tensor cores win (128 taps: 438 against 227)
                const bool cx = f->mma.cplx != 0;                  // complex64: the kernel runs on the float view of x, y, history
                // complex64 (measured, split form): rational 147//160 345 against 311 on k_tiled_c64, 160//147 357 against 286,
                // interpolator 4//1 495 against 431 on k_unit_c64, standard x 64 taps 338 against 165; short standard filters stay
                // on the tiled / unit kernels, decimators on k_decim / k_decim8 (MRB_C64_UNIT_FIRST=1: the unit kernel first)
                static const bool cx_unit_first = getenv("MRB_C64_UNIT_FIRST") != nullptr;
                const bool unit_better = cx ? ((cx_unit_first && f->unit.ok) || (f->unit.ok && f->L == 1 && f->M == 1 && f->T <= 24) ||
                                               (f->kind == MRB_DECIMATOR && f->decim.ok))
                                            : (f->unit.ok && f->L == 1 && f->M == 1 && f->T <= 48);
                if (mma_int && !unit_better) {
                    MmaSched S{};
                    S.mode = 2; S.L = f->L; S.M = f->M; S.p0 = P.p0; S.d0m1 = P.d0m1; S.k_base = 0; S.cplx = cx ? 1 : 0;
                    S.span_c = (15 * f->M) / f->L + 1;
                    GenParams Pv = P;
                    if (cx) { Pv.ldx *= 2; Pv.n_in *= 2; Pv.H *= 2; Pv.ldy *= 2; }
                    // widest spread of window starts in a group: 32 outputs, or 16 complex outputs seen as 32 floats
                    const int64_t span = cx ? 2 * ((15 * f->M) / f->L + 1) + 1 : (31 * f->M) / f->L + 1;
                    k_begin = mma_try_launch(f->mma, f->tctx[ctx].mrows, Pv, S, 0, f->d_bank, nullptr, nullptr, 0, cx ? 2 * N : N,
                                             span, st, &f->last_kernel, &f->launches);
                    if (k_begin == -2) return fail(MRB_ERR_CUDA, "tensor-core launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                }
            }
            // decimators: the chunk head (outputs whose long windows reach the history: 20 us of k_head_warp at C2 beside 174 us
            // of main kernel) runs on the side stream, forked here and joined below
            TableCtx &tci = f->tctx[ctx];
            const bool may_fork = f->policy != 1 && f->kind == MRB_DECIMATOR && f->decim.ok && k_begin == -1;
            if (may_fork) {
                int32_t rc = ensure_side(tci);
                if (rc) return rc;
                if (tci.side) CU(cudaEventRecord(tci.ev_fork, st));
            }
            if (k_begin == -1 && f->policy != 1) {
                k_begin = tiled_try_launch(f->tiled, P, st, &f->last_kernel, &f->launches);
                if (k_begin == -1) k_begin = unit_try_launch(f->unit, P, st, &f->last_kernel, &f->launches);
                if (k_begin == -1) k_begin = decim_try_launch(f->decim, P, st, &f->last_kernel, &f->launches);
                if (k_begin == -2) return fail(MRB_ERR_CUDA, "fast-path launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
            if (k_begin == -1 && f->policy != 1 && f->table.ok && nc >= 32) {
                // Float64 samples: the table kernel's FP64 tensor-core variant with the schedule in closed form (64-channel
                // CTAs: with a handful of channels k_stream wins)
                GenParams Pt = P;
                Pt.sn = nullptr; Pt.sphi = nullptr; Pt.salpha = nullptr;
                // outputs whose window reaches the history: n_k < H  <=>  k < ceil(((H - d0m1) L - p0) / M)
                const int64_t head = std::min<int64_t>(N, std::max<int64_t>(0, ceil_div((f->H - P.d0m1) * f->L - P.p0, f->M)));
                k_begin = table_try_launch(f->table, tci.rows, Pt, f->kind, 1, f->th == MRB_F32, f->d_bank, nullptr, nullptr,
                                           (double)f->L / (double)f->M, 0, N, head, (7 * f->M) / f->L + 1, st, &f->last_kernel,
                                           &f->launches, 0);
                if (k_begin == -2) return fail(MRB_ERR_CUDA, "table launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
            if (k_begin != 0) {            // generic kernel: everything, or the head the tiled kernel left out
                cudaStream_t hs = st;
                if (k_begin > 0) {
                    P.nout = k_begin;
                    if (may_fork && tci.side && k_begin * nc >= 8192) { CU(cudaStreamWaitEvent(tci.side, tci.ev_fork, 0)); hs = tci.side; }
                }
                const char *gname = dispatch_generic(f, P, hs);
                if (k_begin < 0) f->last_kernel = gname;
                ++f->launches;
                if (hs != st) { CU(cudaEventRecord(tci.ev_join, hs)); CU(cudaStreamWaitEvent(st, tci.ev_join, 0)); }
            }
        } else {
            TableCtx &tc = f->tctx[ctx];
            int32_t rc = ensure_sched(f, tc);
            if (rc) return rc;
            // the exact host replay (data independent) was done by check_filt_args; uploaded in bounded sub-chunks
            SchedStore &store = f->sched[f->sched_cur];
            const int64_t *vn = store.n; const int32_t *vphi = store.phi; const double *va = store.a;
            if ((int64_t)store.size != N || !f->sched_valid) return fail(MRB_ERR_BAD_ARGUMENT, "internal: schedule not prepared");
            if (!store.ev) CU(cudaEventCreateWithFlags(&store.ev, cudaEventDisableTiming));
            int si = 0;
            for (int64_t k0 = 0; k0 < N; k0 += kSchedChunk, si ^= 1) {
                const int64_t cnt = std::min(kSchedChunk, N - k0);
                SchedSlot &s = tc.slot[si];
                // uploaded straight from the pinned store (stream order keeps a slot's previous readers ahead of this write)
                // (a later channel block of the same host call on this stream finds slice, tap rows and tiles in place)
                const uint64_t tag = f->serial * 65536 + (uint64_t)(k0 / kSchedChunk) + 1;
                if (s.tag != tag) {
                    CU(cudaMemcpyAsync(s.d_n, vn + k0, cnt * sizeof(int64_t), cudaMemcpyHostToDevice, st));
                    CU(cudaMemcpyAsync(s.d_a, va + k0, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
                    if (f->kind == MRB_ARBITRARY)
                        CU(cudaMemcpyAsync(s.d_phi, vphi + k0, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                    s.tag = tag;
                    if (k0 + cnt >= N) { CU(cudaEventRecord(store.ev, st)); store.pending = true; }   // the store is free again after this
                }
                P.sn = s.d_n; P.k_base = k0; P.nout = cnt;
                P.sphi = s.d_phi; P.salpha = s.d_a;
                cudaStream_t hs = st;                          // stream of the slice's head
                if (f->policy != 1 && tc.side) CU(cudaEventRecord(tc.ev_fork, st));
                if (f->policy != 1) {
                    // fast path: per-output tap rows built once for all channels, then one dot product per output
                    int64_t head = 0;                          // outputs of the slice whose window reaches the history
                    while (head < cnt && vn[k0 + head] < f->H) ++head;
                    int64_t kb = -1;
                    if (f->mma.ok && f->policy == 0) {         // tensor cores first (float32 samples and taps)
                        const bool cx = f->mma.cplx != 0;      // complex64: float view, a group is 16 complex outputs
                        const int64_t og = cx ? kMmaG / 2 : kMmaG;
                        int64_t gs32 = 0;                      // widest spread of window starts inside a group
                        for (int64_t g0 = 0; g0 < cnt; g0 += og)
                            gs32 = std::max(gs32, vn[k0 + std::min<int64_t>(g0 + og, cnt) - 1] - vn[k0 + g0]);
                        MmaSched S{};
                        S.mode = f->kind == MRB_FARROW ? 1 : 0; S.sn = s.d_n; S.sphi = s.d_phi; S.sa = s.d_a; S.cplx = cx ? 1 : 0;
                        GenParams Pv = P;
                        if (cx) { Pv.ldx *= 2; Pv.n_in *= 2; Pv.H *= 2; Pv.ldy *= 2; S.span_c = gs32; gs32 = 2 * gs32 + 1; }
                        kb = mma_try_launch(f->mma, tc.mrows, Pv, S, f->polyorder + 1, f->d_bank, f->d_dbank, f->d_pnfb, cx ? 2 * k0 : k0,
                                            cx ? 2 * cnt : cnt, gs32, st, &f->last_kernel, &f->launches, tag);
                        if (cx && kb > 0) kb /= 2;
                        if (kb == -2) return fail(MRB_ERR_CUDA, "tensor-core launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    }
                    int64_t gspan = 0;                         // widest spread of window starts inside a group of 8 outputs
                    if (kb == -1)
                        for (int64_t g0 = 0; g0 < cnt; g0 += kTabGroup)
                            gspan = std::max(gspan, vn[k0 + std::min(g0 + kTabGroup, cnt) - 1] - vn[k0 + g0]);
                    if (kb == -1) kb = table_try_launch(f->table, tc.rows, P, f->kind, f->polyorder + 1, f->th == MRB_F32, f->d_bank,
                                                        f->d_dbank, f->d_pnfb, f->rate, k0, cnt, head, gspan, st, &f->last_kernel,
                                                        &f->launches, tag);
                    if (kb == -2) return fail(MRB_ERR_CUDA, "table launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    if (kb == 0) continue;
                    if (kb > 0) {                              // the generic kernel computes the slice's head
                        P.nout = kb;
                        if (tc.side) { CU(cudaStreamWaitEvent(tc.side, tc.ev_fork, 0)); hs = tc.side; }
                    } else {
                        f->last_kernel = "generic";
                    }
                } else {
                    f->last_kernel = "generic";
                }
                if (f->kind == MRB_ARBITRARY) {
                    P.mode = SEQ_ARBITRARY;
                } else {
                    P.mode = SEQ_FARROW; P.taptab = tc.d_taptab;
                    // tap rows for the outputs the generic kernel computes: the whole slice, or only its head
                    const unsigned g = (unsigned)ceil_div(P.nout * f->T, 256);
                    if (is_double(f->ty))
                        k_farrow_taps<double><<<g, 256, 0, hs>>>(f->d_pnfb, f->polyorder + 1, f->T, s.d_a, P.nout, (double *)tc.d_taptab, f->th == MRB_F32);
                    else
                        k_farrow_taps<float><<<g, 256, 0, hs>>>(f->d_pnfb, f->polyorder + 1, f->T, s.d_a, P.nout, (float *)tc.d_taptab, f->th == MRB_F32);
                    ++f->launches;
                }
                dispatch_generic(f, P, hs);
                ++f->launches;
                if (hs != st) { CU(cudaEventRecord(tc.ev_join, hs)); CU(cudaStreamWaitEvent(st, tc.ev_join, 0)); }
            }
        }
    }
    if (t0) { CU(cudaEventRecord(t1, st)); f->tev.emplace_back(t0, t1); }
    if (n_in > 0 && f->H > 0) {
        launch_history(f, x, ldx, n_in, hold, hnew, nc, st);
        ++f->launches;
    }
    CU(cudaGetLastError());
    return MRB_OK;
}

static int32_t check_filt_args(mrb_filter *f, const void *x, int64_t ldx, int64_t n_in, void *y, int64_t ldy,
                               int64_t cap, int64_t *N, mrb_state *end) {
    if (!f) return fail(MRB_ERR_BAD_ARGUMENT, "null handle");
    if (f->device < 0) return fail(MRB_ERR_NO_DEVICE, "host-only handle: no CUDA device bound, and there is no CPU fallback");
    if (n_in < 0 || (n_in > 0 && !x) || ldx < n_in) return fail(MRB_ERR_BAD_ARGUMENT, "bad x / ld_x / n_in");
    *N = count_outputs(f, n_in, end);          // table kinds: fills (or reuses) the cached schedule store
    if (*N < 0) return fail(MRB_ERR_CUDA, "out of memory for the schedule");
    if (*N > cap) {
        const char *msg = f->kind == MRB_STANDARD ? "buffer length must be >= x length"                    // :460
                          : f->kind == MRB_INTERPOLATOR ? "length( buffer ) must be >= interpolation * length(x)"  // :503
                                                        : "buffer is too small";                                   // :550
        return fail(MRB_ERR_BUFFER_TOO_SMALL, "%s", msg);
    }
    if (*N > 0 && (!y || ldy < *N)) return fail(MRB_ERR_BAD_ARGUMENT, "bad y / ld_y");
    return MRB_OK;
}

extern "C" int32_t mrb_filt(mrb_filter *f, const void *x, int64_t ldx, int64_t n_in, void *y, int64_t ldy,
                            int64_t cap, int64_t *n_out, void *stream) {
    int64_t N;
    mrb_state end;
    int32_t rc = check_filt_args(f, x, ldx, n_in, y, ldy, cap, &N, &end);
    if (rc) return rc;
    ++f->serial;
    DeviceGuard guard(f->device);
    rc = order_after_last(f, (cudaStream_t)stream);
    if (rc) return rc;
    rc = run_channels(f, x, ldx, n_in, y, ldy, N, 0, f->nch, (cudaStream_t)stream, 0);
    if (rc) return rc;
    f->last_stream = (cudaStream_t)stream; f->last_valid = true;
    commit_state(f, end);
    if (n_in > 0 && f->H > 0) f->cur ^= 1;
    if (n_out) *n_out = N;
    return MRB_OK;
}

static int32_t grow(void **p, size_t *have, size_t want) {
    if (*have >= want) return MRB_OK;
    cudaFree(*p);
    *p = nullptr; *have = 0;
    CU(cudaMalloc(p, want));
    *have = want;
    return MRB_OK;
}

// Host-buffer form: channel blocks are pipelined H2D -> kernels -> D2H on two streams so the copies of
// one block overlap the compute of the other (PCIe is full duplex).
extern "C" int32_t mrb_filt_host(mrb_filter *f, const void *x, int64_t ldx, int64_t n_in, void *y, int64_t ldy,
                                 int64_t cap, int64_t *n_out) {
    int64_t N;
    mrb_state end;
    int32_t rc = check_filt_args(f, x, ldx, n_in, y, ldy, cap, &N, &end);
    if (rc) return rc;
    ++f->serial;
    DeviceGuard guard(f->device);
    const size_t es = dsize(f->tx), eo = dsize(f->ty);
    // channel blocks of ~MRB_HOST_BLOCK_MIB (64) MiB of input, pipelined over MRB_HOST_STREAMS (1..4, default 2) streams
    // (measured: 16..256 MiB x 2..4 streams all land on the same 86 GB/s full-duplex PCIe limit)
    // Default block: 64 MiB, but at least ~16 blocks per call -- the first block's upload and the last block's download have
    // nothing to overlap with, which costs 1/(blocks) of the call (a 0.25 GiB call in four blocks ran at 64 % of the copy ceiling)
    static const int env_mib = getenv("MRB_HOST_BLOCK_MIB") ? std::max(1, atoi(getenv("MRB_HOST_BLOCK_MIB"))) : 0;
    static const int env_nst = getenv("MRB_HOST_STREAMS") ? std::min(kMaxHostStreams, std::max(1, atoi(getenv("MRB_HOST_STREAMS")))) : 2;
    const int64_t total_mib = (int64_t)(((size_t)f->nch * (size_t)std::max<int64_t>(n_in, 1) * es) >> 20);
    const int blk_mib = f->host_block_mib > 0 ? f->host_block_mib : env_mib > 0 ? env_mib : (int)std::min<int64_t>(64, std::max<int64_t>(8, total_mib / 16));
    const int nst = f->host_streams_n > 0 ? f->host_streams_n : env_nst;
    int64_t cb = std::max<int64_t>(1, (int64_t)(((size_t)blk_mib << 20) / std::max<size_t>(1, (size_t)n_in * es)));
    cb = std::min(cb, f->nch);
    // staging row pitches are multiples of 16 bytes, so the TMA fast paths apply whatever n_in and N are
    const int64_t ax = std::max<int64_t>(1, 16 / (int64_t)es), ay = std::max<int64_t>(1, 16 / (int64_t)eo);
    const int64_t lxs = (std::max<int64_t>(n_in, 1) + ax - 1) / ax * ax, lys = (std::max<int64_t>(N, 1) + ay - 1) / ay * ay;
    const size_t xb = (size_t)cb * lxs * es, yb = (size_t)cb * lys * eo;
    rc = grow(&f->d_xs, &f->xs_bytes, nst * xb); if (rc) return rc;
    rc = grow(&f->d_ys, &f->ys_bytes, nst * yb); if (rc) return rc;
    cudaStream_t sts[kMaxHostStreams] = {f->own_stream};
    for (int i = 1; i < nst; ++i) {                        // the handle's own streams, on the handle's device
        if (!f->host_streams[i - 1]) CU(cudaStreamCreateWithFlags(&f->host_streams[i - 1], cudaStreamNonBlocking));
        sts[i] = f->host_streams[i - 1];
    }
    for (int i = 0; i < nst; ++i) { rc = order_after_last(f, sts[i]); if (rc) return rc; }
    int b = 0;
    for (int64_t c0 = 0; c0 < f->nch; c0 += cb, b = (b + 1) % nst) {
        const int64_t nc = std::min(cb, f->nch - c0);
        char *dx = static_cast<char *>(f->d_xs) + b * xb, *dy = static_cast<char *>(f->d_ys) + b * yb;
        cudaStream_t st = sts[b];
        // (contiguous rows on both sides -- the usual case -- go as ONE 1-D copy: the DMA engines do fewer, longer bursts)
        if (n_in > 0) {
            if (ldx == n_in && lxs == n_in)
                CU(cudaMemcpyAsync(dx, static_cast<const char *>(x) + (size_t)(c0 * ldx) * es, (size_t)nc * (size_t)n_in * es, cudaMemcpyHostToDevice, st));
            else
                CU(cudaMemcpy2DAsync(dx, (size_t)lxs * es, static_cast<const char *>(x) + (size_t)(c0 * ldx) * es,
                                     (size_t)ldx * es, (size_t)n_in * es, (size_t)nc, cudaMemcpyHostToDevice, st));
        }
        // arbitrary / farrow: every block uploads the schedule slices and builds the tap rows it needs into the table
        // context of ITS stream (TableCtx): blocks in flight on other streams never see these buffers change
        rc = run_channels(f, dx, lxs, n_in, dy, lys, N, c0, nc, st, b);
        if (rc) return rc;
        if (N > 0) {
            if (ldy == N && lys == N)                      // contiguous on both sides (only the N valid samples of a row are ever written)
                CU(cudaMemcpyAsync(static_cast<char *>(y) + (size_t)(c0 * ldy) * eo, dy, (size_t)nc * (size_t)N * eo, cudaMemcpyDeviceToHost, st));
            else
                CU(cudaMemcpy2DAsync(static_cast<char *>(y) + (size_t)(c0 * ldy) * eo, (size_t)ldy * eo, dy, (size_t)lys * eo,
                                     (size_t)N * eo, (size_t)nc, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int i = 0; i < nst; ++i) CU(cudaStreamSynchronize(sts[i]));
    f->last_valid = false;                                 // nothing of this handle is in flight any more
    for (auto &sst : f->sched) sst.pending = false;        // (every stream that read a schedule store has been synchronised)
    commit_state(f, end);
    if (n_in > 0 && f->H > 0) f->cur ^= 1;
    if (n_out) *n_out = N;
    return MRB_OK;
}

// ------------------------------------------------------------------------------------------
// segment split
// ------------------------------------------------------------------------------------------
template <typename RX, int NC>
__global__ void k_load_halo(const RX *__restrict__ halo, int64_t ld, RX *__restrict__ hist, int64_t H, int64_t nch) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= H * nch) return;
    const int64_t c = idx / H, i = idx - c * H;
    RX s[NC];
    ld_sample<RX, NC>(halo + c * ld * NC, i, s);
    st_sample<RX, NC>(hist + c * H * NC, i, s);
}

extern "C" int32_t mrb_seek(mrb_filter *f, int64_t n0, const void *halo, int64_t ld_halo, int64_t *k0, void *stream) {
    if (!f || n0 < 0) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    int64_t kk;
    if (is_table_kind(f)) {
        // arbitrary / farrow: the Float64 phase recurrence has no closed form that is bit-identical to the reference's
        // rounding, so the start state is the exact (data independent) replay of n0 inputs from the constructor state
        // (src/Filters.jl:663-673, 780-786), without storing the schedule: ~2 ns per output on the host.
        init_state(f);
        mrb_state e;
        kk = replay_table(f, n0, &e, nullptr);
        commit_state(f, e);
        f->sched_valid = false;
    } else {
        // state after consuming n0 samples from the constructor state (p=0, d=1): SURVEY 8e
        kk = ceil_div(n0 * f->L, f->M);
        f->phiIdx = (kk * f->M) % f->L + 1;
        f->deficit = (kk * f->M) / f->L - n0 + 1;
    }
    if (k0) *k0 = kk;
    if (f->device >= 0 && f->H > 0) {
        DeviceGuard guard(f->device);
        cudaStream_t st = (cudaStream_t)stream;
        {   int32_t rc = order_after_last(f, st); if (rc) return rc; }
        f->last_stream = st; f->last_valid = true;
        void *dst = f->d_hist[f->cur];
        if (!halo) {
            CU(cudaMemsetAsync(dst, 0, (size_t)(f->H * f->nch) * dsize(f->tx), st));
        } else {
            if (ld_halo < f->H) return fail(MRB_ERR_BAD_ARGUMENT, "ld_halo < history_len");
            const unsigned g = (unsigned)ceil_div(f->H * f->nch, 256);
            switch (f->tx) {
            case MRB_F32: k_load_halo<float, 1><<<g, 256, 0, st>>>((const float *)halo, ld_halo, (float *)dst, f->H, f->nch); break;
            case MRB_F64: k_load_halo<double, 1><<<g, 256, 0, st>>>((const double *)halo, ld_halo, (double *)dst, f->H, f->nch); break;
            case MRB_C64: k_load_halo<float, 2><<<g, 256, 0, st>>>((const float *)halo, ld_halo, (float *)dst, f->H, f->nch); break;
            case MRB_C128: k_load_halo<double, 2><<<g, 256, 0, st>>>((const double *)halo, ld_halo, (double *)dst, f->H, f->nch); break;
            }
            ++f->launches;
            CU(cudaGetLastError());
        }
    }
    return MRB_OK;
}

extern "C" int32_t mrb_launch_count(const mrb_filter *f, int64_t *n) {
    if (!f || !n) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    *n = f->launches;
    return MRB_OK;
}

extern "C" int32_t mrb_set_kernel_policy(mrb_filter *f, int32_t policy) {
    if (!f || policy < 0 || policy > 2) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    f->policy = policy;
    return MRB_OK;
}

extern "C" int32_t mrb_set_host_pipeline(mrb_filter *f, int32_t block_mib, int32_t n_streams) {
    if (!f || block_mib < 0 || n_streams < 0 || n_streams > kMaxHostStreams) return fail(MRB_ERR_BAD_ARGUMENT, "bad argument");
    f->host_block_mib = block_mib;
    f->host_streams_n = n_streams;
    return MRB_OK;
}

extern "C" int32_t mrb_set_timing(mrb_filter *f, int32_t on) {
    if (!f) return fail(MRB_ERR_BAD_ARGUMENT, "null handle");
    f->timing = on != 0;
    return MRB_OK;
}

extern "C" int32_t mrb_get_timing(mrb_filter *f, double *mean_ms, int64_t *n_calls) {
    if (!f || !mean_ms) return fail(MRB_ERR_BAD_ARGUMENT, "null argument");
    double sum = 0.0;
    for (auto &p : f->tev) {
        CU(cudaEventSynchronize(p.second));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, p.first, p.second));
        sum += ms;
        cudaEventDestroy(p.first); cudaEventDestroy(p.second);
    }
    *mean_ms = f->tev.empty() ? 0.0 : sum / (double)f->tev.size();
    if (n_calls) *n_calls = (int64_t)f->tev.size();
    f->tev.clear();
    return MRB_OK;
}

extern "C" const char *mrb_last_kernel(const mrb_filter *f) { return f ? f->last_kernel : "none"; }
extern "C" const char *mrb_last_error(void) { return g_err.c_str(); }
extern "C" const char *mrb_version(void) { return "libmrb 0.1.0 (sm_100a)"; }
