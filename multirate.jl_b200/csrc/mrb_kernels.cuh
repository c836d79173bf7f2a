// mrb_kernels.cuh -- generic (any kind, any dtype, any alignment) sm_100a kernels:
//   k_generic      one thread per output sample, loops over its channel slice
//   k_stream       the same for integer schedules, polyphase bank staged in shared memory
//   k_head_warp    chunk heads of long filters: one warp per output and channel
//   k_farrow_taps  per-output Farrow tap rows (Float64 Horner, rounded to the tap type)
//   k_history      history carry  hist <- last H of [hist | x]   (shiftin!, src/support.jl:61-80)
// The tiled fast paths live in mrb_tiled.cuh; this file is the always-correct path that
// every (kind, dtype) combination can fall back to ON THE GPU (there is no CPU path).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mrb {

template <typename R> struct Vec2T;
template <> struct Vec2T<float> { using type = float2; };
template <> struct Vec2T<double> { using type = double2; };

template <typename RX, int NC>
__device__ __forceinline__ void ld_sample(const RX *__restrict__ base, int64_t idx, RX (&out)[NC]) {
    if constexpr (NC == 1) {
        out[0] = __ldg(base + idx);
    } else {
        using V = typename Vec2T<RX>::type;
        const V t = __ldg(reinterpret_cast<const V *>(base) + idx);
        out[0] = t.x;
        out[1] = t.y;
    }
}

template <typename R, int NC>
__device__ __forceinline__ void st_sample(R *__restrict__ base, int64_t idx, const R (&v)[NC]) {
    if constexpr (NC == 1) {
        base[idx] = v[0];
    } else {
        using V = typename Vec2T<R>::type;
        V t;
        t.x = v[0];
        t.y = v[1];
        reinterpret_cast<V *>(base)[idx] = t;
    }
}

enum { SEQ_INTEGER = 0, SEQ_ARBITRARY = 1, SEQ_FARROW = 2 };

struct GenParams {
    const void *x;      // channel-major samples, first channel of this slice
    int64_t ldx, n_in;
    const void *hist;   // [nch][H]
    int64_t H;
    void *y;
    int64_t ldy;
    const void *bank;   // R[Nphi][T]: row phi = reference pfb[:, phi] (time reversed branch)
    const void *dbank;  // arbitrary only
    int64_t T;
    int32_t mode;
    int64_t L, M, p0, d0m1;     // integer schedule: n_k = d0m1 + (p0 + k*M)/L, phi_k = (p0 + k*M) % L
    const int64_t *sn;          // table schedules: 0-based x index of the last window sample
    const int32_t *sphi;        // arbitrary: 0-based branch
    const double *salpha;       // arbitrary: alpha
    const void *taptab;         // farrow: R[nout][T]
    int64_t k_base;             // first output of this launch (index into y and, for SEQ_INTEGER, into the schedule)
    int64_t nout;               // outputs in this launch
    int64_t nch;
    int32_t cpb_log2;           // k_generic: a block is (256 >> cpb_log2) outputs x (1 << cpb_log2) channels
};

// y[c, k] = sum_i taps_k[i] * ext[c, n_k + i],  ext = [hist | x]  (H = T-1, so the window of the output
// whose last sample is x[n_k] starts at ext index n_k).  Reference: src/support.jl:5-55 called from
// src/Filters.jl:462-468, 505-512, 558-569, 613-625, 717-732, 814-826.
// Block shape: 256 outputs of one channel, or -- for the short launches that compute a chunk's head (the few outputs whose
// window reaches the history) -- 256 >> s outputs of 1 << s channels, so that the block's threads all have work.
template <typename RX, typename R, int NC>
__global__ void __launch_bounds__(256) k_generic(const GenParams P) {
    const int W = 256 >> P.cpb_log2;
    const int64_t kl = (int64_t)blockIdx.x * W + (threadIdx.x & (W - 1));
    const int cl = threadIdx.x >> (8 - P.cpb_log2), cpb = 1 << P.cpb_log2;
    if (kl >= P.nout) return;
    const int64_t k = P.k_base + kl;
    const R *__restrict__ taps;
    const R *__restrict__ dtaps = nullptr;
    int64_t n;
    double alpha = 0.0;
    if (P.mode == SEQ_INTEGER) {
        const int64_t t = P.p0 + k * P.M;
        int64_t tq, tr;
        if ((uint64_t)t < (1ull << 32) && (uint64_t)P.L < (1ull << 32)) {   // the usual case: one 32-bit division
            tq = (uint32_t)t / (uint32_t)P.L;
            tr = (uint32_t)t - (uint32_t)tq * (uint32_t)P.L;
        } else {
            tq = t / P.L;
            tr = t - tq * P.L;
        }
        n = P.d0m1 + tq;
        taps = static_cast<const R *>(P.bank) + tr * P.T;
    } else if (P.mode == SEQ_ARBITRARY) {
        n = P.sn[kl];
        const int64_t phi = P.sphi[kl];
        taps = static_cast<const R *>(P.bank) + phi * P.T;
        dtaps = static_cast<const R *>(P.dbank) + phi * P.T;
        alpha = P.salpha[kl];
    } else {
        n = P.sn[kl];
        taps = static_cast<const R *>(P.taptab) + kl * P.T;
    }
    const int64_t H = P.H;
    const int T = (int)P.T;
    // taps i in [0, ih) read history, [ih, T) read x
    int64_t ihl = H - n;
    const int ih = (int)(ihl < 0 ? 0 : (ihl > T ? T : ihl));
    for (int64_t c = (int64_t)blockIdx.y * cpb + cl; c < P.nch; c += (int64_t)gridDim.y * cpb) {
        const RX *__restrict__ xc = static_cast<const RX *>(P.x) + c * P.ldx * NC;
        const RX *__restrict__ hc = static_cast<const RX *>(P.hist) + c * H * NC;
        R acc[NC], dacc[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = dacc[q] = R(0);
#pragma unroll 4
        for (int i = 0; i < ih; ++i) {
            RX s[NC];
            ld_sample<RX, NC>(hc, n + i, s);
            const R t = __ldg(taps + i);
#pragma unroll
            for (int q = 0; q < NC; ++q) acc[q] = fma(t, (R)s[q], acc[q]);
            if (dtaps) {
                const R dt = __ldg(dtaps + i);
#pragma unroll
                for (int q = 0; q < NC; ++q) dacc[q] = fma(dt, (R)s[q], dacc[q]);
            }
        }
        const RX *__restrict__ xw = xc + (n - H) * NC;
        if (dtaps) {
#pragma unroll 4
            for (int i = ih; i < T; ++i) {
                RX s[NC];
                ld_sample<RX, NC>(xw, i, s);
                const R t = __ldg(taps + i), dt = __ldg(dtaps + i);
#pragma unroll
                for (int q = 0; q < NC; ++q) {
                    acc[q] = fma(t, (R)s[q], acc[q]);
                    dacc[q] = fma(dt, (R)s[q], dacc[q]);
                }
            }
            // yLower + yUpper*alpha with Float64 alpha (src/Filters.jl:730)
#pragma unroll
            for (int q = 0; q < NC; ++q) acc[q] = (R)((double)acc[q] + (double)dacc[q] * alpha);
        } else {
#pragma unroll 4
            for (int i = ih; i < T; ++i) {
                RX s[NC];
                ld_sample<RX, NC>(xw, i, s);
                const R t = __ldg(taps + i);
#pragma unroll
                for (int q = 0; q < NC; ++q) acc[q] = fma(t, (R)s[q], acc[q]);
            }
        }
        st_sample<R, NC>(static_cast<R *>(P.y) + c * P.ldy * NC, k, acc);
    }
}

// Chunk heads of long filters (every schedule kind): the few outputs whose window reaches the history, for every channel.
// With one thread per output a warp's loads are T-strided and each lane walks its whole window alone (decimator 1//8 x
// 256 taps, 1024 channels: 31 us for 40 outputs per channel, 14 % of the step).  Here a WARP computes one output of one
// channel: lane i takes taps i, i+32, ... (consecutive lanes read consecutive samples and taps), then a shuffle tree.
template <typename RX, typename R, int NC>
__global__ void __launch_bounds__(256) k_head_warp(const GenParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;            // (channel, output) pair of this warp
    if (w >= P.nout * P.nch) return;
    const int64_t c = w / P.nout, kl = w - c * P.nout;
    const int64_t k = P.k_base + kl;
    const int64_t H = P.H;
    const int T = (int)P.T;
    int64_t n;
    const R *__restrict__ taps;
    const R *__restrict__ dtaps = nullptr;
    double alpha = 0.0;
    if (P.mode == SEQ_INTEGER) {
        const int64_t t = P.p0 + k * P.M;
        const int64_t tq = t / P.L, tr = t - tq * P.L;
        n = P.d0m1 + tq;
        taps = static_cast<const R *>(P.bank) + tr * P.T;
    } else if (P.mode == SEQ_ARBITRARY) {                                        // as k_generic: two dots, Float64 blend (:723-730)
        n = P.sn[kl];
        const int64_t phi = P.sphi[kl];
        taps = static_cast<const R *>(P.bank) + phi * P.T;
        dtaps = static_cast<const R *>(P.dbank) + phi * P.T;
        alpha = P.salpha[kl];
    } else {
        n = P.sn[kl];
        taps = static_cast<const R *>(P.taptab) + kl * P.T;
    }
    const RX *__restrict__ hc = static_cast<const RX *>(P.hist) + c * H * NC;
    const RX *__restrict__ xw = static_cast<const RX *>(P.x) + (c * P.ldx + (n - H)) * NC;
    R acc[NC], dacc[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) acc[q] = dacc[q] = R(0);
#pragma unroll 4
    for (int i = lane; i < T; i += 32) {
        RX s[NC];
        if (n + i < H) ld_sample<RX, NC>(hc, n + i, s);
        else ld_sample<RX, NC>(xw, i, s);
        const R tv = __ldg(taps + i);
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = fma(tv, (R)s[q], acc[q]);
        if (dtaps) {
            const R dv = __ldg(dtaps + i);
#pragma unroll
            for (int q = 0; q < NC; ++q) dacc[q] = fma(dv, (R)s[q], dacc[q]);
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            dacc[q] += __shfl_xor_sync(0xffffffffu, dacc[q], o);
        }
    if (dtaps) {
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = (R)((double)acc[q] + (double)dacc[q] * alpha);
    }
    if (lane == 0) st_sample<R, NC>(static_cast<R *>(P.y) + c * P.ldy * NC, k, acc);
}

// Integer schedules (standard / interpolator / decimator / rational) that no tiled kernel covers -- few channels,
// Float64, odd alignments: k_generic's thread-per-output mapping, but with the polyphase bank staged in shared memory.
// Consecutive outputs use different branches (phi advances by M mod L), so a warp's tap loads touch 32 different rows:
// from global memory that is 32 sectors per request and the L1 data path bounds the kernel (ncu, README benchmark:
// 18.6 sectors per request, 39 us for 918,750 outputs); from shared memory with an odd row pitch the 32 rows fall
// into 32 different banks.  Persistent grid-stride CTAs so the bank is staged once per CTA, not once per 256 outputs.
template <typename RX, typename R, int NC>
__global__ void __launch_bounds__(256) k_stream(const GenParams P, int pitch) {
    extern __shared__ __align__(16) unsigned char stream_smem[];
    R *sb = reinterpret_cast<R *>(stream_smem);
    const int T = (int)P.T;
    {
        const R *__restrict__ gb = static_cast<const R *>(P.bank);
        const int total = (int)P.L * T;
        for (int idx = threadIdx.x; idx < total; idx += 256) {
            const int phi = idx / T;
            sb[phi * pitch + (idx - phi * T)] = __ldg(gb + idx);
        }
    }
    __syncthreads();
    const int64_t H = P.H;
    for (int64_t kl = (int64_t)blockIdx.x * 256 + threadIdx.x; kl < P.nout; kl += (int64_t)gridDim.x * 256) {
        const int64_t k = P.k_base + kl;
        const int64_t t = P.p0 + k * P.M;
        int64_t tq, tr;
        if ((uint64_t)t < (1ull << 32)) {
            tq = (uint32_t)t / (uint32_t)P.L;
            tr = (uint32_t)t - (uint32_t)tq * (uint32_t)P.L;
        } else {
            tq = t / P.L;
            tr = t - tq * P.L;
        }
        const int64_t n = P.d0m1 + tq;
        const R *taps = sb + (int)tr * pitch;
        const int64_t ihl = H - n;
        const int ih = (int)(ihl < 0 ? 0 : (ihl > T ? T : ihl));
        for (int64_t c = blockIdx.y; c < P.nch; c += gridDim.y) {
            const RX *__restrict__ hc = static_cast<const RX *>(P.hist) + c * H * NC;
            const RX *__restrict__ xw = static_cast<const RX *>(P.x) + (c * P.ldx + (n - H)) * NC;
            R acc[NC];
#pragma unroll
            for (int q = 0; q < NC; ++q) acc[q] = R(0);
            for (int i = 0; i < ih; ++i) {
                RX s[NC];
                ld_sample<RX, NC>(hc, n + i, s);
#pragma unroll
                for (int q = 0; q < NC; ++q) acc[q] = fma(taps[i], (R)s[q], acc[q]);
            }
#pragma unroll 8
            for (int i = ih; i < T; ++i) {
                RX s[NC];
                ld_sample<RX, NC>(xw, i, s);
#pragma unroll
                for (int q = 0; q < NC; ++q) acc[q] = fma(taps[i], (R)s[q], acc[q]);
            }
            st_sample<R, NC>(static_cast<R *>(P.y) + c * P.ldy * NC, k, acc);
        }
    }
}

// currentTaps[i] = polyval(pnfb[i], phiIdx) for every output of the launch (src/Filters.jl:789-791):
// Horner highest order first in Float64 with separately rounded multiply and add (no FMA contraction,
// as Polynomials.polyval), then rounded to the tap type.
template <typename R>
__global__ void __launch_bounds__(256) k_farrow_taps(const double *__restrict__ pnfb, int P1, int64_t T,
                                                     const double *__restrict__ phase, int64_t nout,
                                                     R *__restrict__ taptab, int tap_is_f32) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= nout * T) return;
    const int64_t k = idx / T, i = idx - k * T;
    const double ph = phase[k];
    const double *c = pnfb + i * P1;
    double v = c[P1 - 1];
    for (int p = P1 - 2; p >= 0; --p) v = __dadd_rn(__dmul_rn(v, ph), c[p]);
    if (tap_is_f32) v = (double)(float)v;
    taptab[idx] = (R)v;
}

// hist_new[c][i] = ext[c][n_in + i], ext = [hist_old | x]   (shiftin!, src/support.jl:61-80).
// Always double buffered: when n_in < H source and destination overlap.
template <typename RX, int NC>
__global__ void __launch_bounds__(256) k_history(const RX *__restrict__ x, int64_t ldx, int64_t n_in,
                                                 const RX *__restrict__ hold, RX *__restrict__ hnew, int64_t H,
                                                 int64_t nch) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= H * nch) return;
    const int64_t c = idx / H, i = idx - c * H;
    const int64_t j = i + n_in;
    RX s[NC];
    if (j < H) ld_sample<RX, NC>(hold + c * H * NC, j, s);
    else ld_sample<RX, NC>(x + c * ldx * NC, j - H, s);
    st_sample<RX, NC>(hnew + c * H * NC, i, s);
}

}  // namespace mrb
