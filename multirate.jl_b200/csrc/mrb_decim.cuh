// mrb_decim.cuh -- fast path for FIRDecimator (src/Filters.jl:598-631) on complex64 or float32 samples x float32 taps
// (BASELINE configs[1]: 1//8, 256 taps, 1024 channels, streamed 64K-sample chunks).
//
// y[k] = sum_i hflip[i] x[kM + i + e]: every output needs M new samples and T old ones, so a lane-per-channel
// mapping would need > 2 KB of shared memory per lane.  Instead the filter is split into its M polyphase
// residues: with i = jM + q,
//     y[k] = sum_q  sum_j hflip[jM + q] * x_q[k + j],      x_q[m] = x[mM + q + e]
// i.e. M unit-stride sub-filters of T/M taps over the decimated streams x_q.  One LANE is one (channel, residue)
// pair: a warp covers 32/M channels, the taps of a lane are M-strided constants held in registers for the whole
// kernel, each lane accumulates M consecutive partial outputs from a sliding register window (plain FFMA, full
// FP32 rate), and a log2(M)-round shuffle reduce-scatter leaves output k0+q in residue lane q -- so a channel's M
// results are written as one contiguous 8M-byte segment.  All 32 lanes of a warp do useful FMAs and a channel costs
// 1/M-th of a warp's shared memory: 16 warps per SM fit.
//  * rows of M samples per channel arrive by TMA ([32 ch][M samples] boxes, one per decimated index m) in a ring
//    laid out [row][channel][M samples]: a warp's LDS.64 reads 256 contiguous bytes, conflict free.
//  * a ninth warp is the TMA producer; consumer warps hand ring slots back through "empty" mbarriers and never meet
//    at a CTA barrier, so one warp's window loads (40 LDS per step) overlap another's FMAs.
// The first outputs of a chunk (window reaches into the history) are computed by k_generic.
#pragma once
#include <cstdio>

#include "mrb_tiled.cuh"

namespace mrb {

constexpr int kDecRows = 32;            // channels per CTA (16-channel CTAs, 4 per SM, measured 6 % slower)
constexpr int kDecTQ = 33;              // tap slots per residue: T <= 32 M, plus one slot of slack for alignment
constexpr int kDecWarps = 8;

struct alignas(16) DecParams {
    long long k_begin, N;      // this launch covers outputs [k_begin, N)
    long long e;               // x index of the first sample of output 0's (padded) window
    int KT;                    // outputs per tile (multiple of M)
    int delta;                 // which tap table: TMA box starts must be 16-byte aligned (an even complex64 sample, a
                               // float32 sample index that is a multiple of 4), so the padded window starts at an
                               // aligned e and the taps are placed delta slots later
    const float *taps;         // device: taps[delta][j * 32 + q] = padded hflip[j*M + q] (tap-major: a warp's loads of
                               // one slot touch M consecutive floats)
};

template <int M, bool CPLX>
struct DecCfg {
    static constexpr int ES = CPLX ? 8 : 4;                         // bytes per sample
    static constexpr int G = 8 / M;                                 // groups of M outputs per lane per step (R = 8)
    static constexpr int R = G * M;                                 // outputs per channel per step
    static constexpr int CPW = 32 / M;                              // channels per warp
    static constexpr int WARPS = kDecRows / CPW;                    // warps per CTA (32 channels)
    static constexpr int ROW_BYTES = kDecRows * M * ES;             // one decimated index, all channels
    static constexpr int LIVE = R + kDecTQ - 1;                     // rows a step reads
    static constexpr int LIVE_SLOTS = (LIVE + R - 1) / R;           // ... in units of R rows (one mbarrier each)
    static constexpr int NSLOT = LIVE_SLOTS + 2;                    // ring: live + two steps ahead
    static constexpr int NROW = NSLOT * R;
    static constexpr int SMEM = NROW * ROW_BYTES + 16 * NSLOT;       // + a full and an empty mbarrier per slot
};

template <int M, bool CPLX>
__global__ void __launch_bounds__(32 * (DecCfg<M, CPLX>::WARPS + 1), 2)
k_decim(const __grid_constant__ CUtensorMap tmx, void *__restrict__ yv, long long ldy, int nch,
        const __grid_constant__ DecParams P) {
    using C = DecCfg<M, CPLX>;
    constexpr int R = C::R, NSLOT = C::NSLOT;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + C::NROW * C::ROW_BYTES);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int q = lane & (M - 1);                                    // residue of this lane
    const int cl = warp * C::CPW + lane / M;                         // channel within the CTA
    const int ch0 = blockIdx.y * kDecRows;
    const uint32_t in_base = smem_u32(smem), bar_base = smem_u32(bars);
    const uint32_t lanepart = (uint32_t)(cl * M + q) * (uint32_t)C::ES;   // this lane's sample inside a row

    const long long k0 = P.k_begin + (long long)blockIdx.x * P.KT;   // first output of the tile
    const int ntile = (int)min((long long)P.KT, P.N - k0);
    const int nsteps = (ntile + R - 1) / R;
    const long long x0 = P.e + k0 * M;                               // x index of row 0 of the tile (even)
    const int glast = nsteps + (C::LIVE - 1) / R;                    // newest row group (R rows) the tile reads

    // full[slot]: the group's R TMA boxes have landed; empty[slot]: every consumer warp has read the group
    const uint32_t empty_base = bar_base + 8 * NSLOT;
    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(bar_base + 8 * i, 1);
            mbar_init(empty_base + 8 * i, C::WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
    }
    __syncthreads();

    if (warp == C::WARPS) {
        // ---- producer warp: one mbarrier per group of R rows, a group is R TMA boxes [32 ch][M samples].  It runs
        // up to NSLOT groups ahead of the slowest consumer warp; the consumer warps never meet at a CTA barrier, so
        // their window loads and FMA phases overlap.
        if (lane == 0) {
            int slot = 0;
            uint32_t par = 0;                                        // round n >= 1 of a slot waits for phase n-1 of its empty barrier
            for (int g = 0; g <= glast; ++g) {
                if (g >= NSLOT) mbar_wait(empty_base + 8 * slot, par);
                const uint32_t bar = bar_base + 8 * slot;
                mbar_expect_tx(bar, R * C::ROW_BYTES);
                for (int r = 0; r < R; ++r)
                    tma_load_2d(in_base + (uint32_t)((slot * R + r) * C::ROW_BYTES), &tmx,
                                (int)((x0 + ((long long)g * R + r) * M) * (CPLX ? 2 : 1)), ch0, bar);
                if (++slot == NSLOT) { slot = 0; if (g >= NSLOT) par ^= 1u; }
            }
        }
        return;
    }

    // this lane's taps: constants for the whole kernel
    float t[kDecTQ];
#pragma unroll
    for (int j = 0; j < kDecTQ; ++j) t[j] = __ldg(P.taps + (P.delta * kDecTQ + j) * 32 + q);

    int g_waited = 0, w_slot = 0;
    uint32_t w_par = 0;

    for (int s = 0; s < nsteps; ++s) {
        // rows [sR, sR + R + TQ - 1) of the tile = groups s .. s + (LIVE-1)/R
        const int need = s + (C::LIVE - 1) / R;
        for (; g_waited <= need; ++g_waited) {
            mbar_wait(bar_base + 8 * w_slot, w_par);
            if (++w_slot == NSLOT) { w_slot = 0; w_par ^= 1u; }
        }
        // complex samples travel as packed (re, im) pairs: one FFMA2 per tap with the lane's tap as its scalar operand
        unsigned long long w2[CPLX ? C::LIVE : 1];
        float w1[CPLX ? 1 : C::LIVE];
        {
            int row = (s % NSLOT) * R;
#pragma unroll
            for (int m = 0; m < C::LIVE; ++m) {
                const uint32_t a = in_base + (uint32_t)(row * C::ROW_BYTES) + lanepart;
                if constexpr (CPLX) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(w2[m]) : "r"(a) : "memory");
                else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w1[m]) : "r"(a) : "memory");
                if (++row == C::NROW) row = 0;
            }
        }
        float2 acc[R];
        if constexpr (CPLX) {
            unsigned long long a2[R];
#pragma unroll
            for (int r = 0; r < R; ++r) a2[r] = 0ull;
#pragma unroll
            for (int j = 0; j < kDecTQ; ++j) {
#pragma unroll
                for (int r = 0; r < R; ++r) cfma(a2[r], t[j], w2[r + j]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[r].x), "=f"(acc[r].y) : "l"(a2[r]));
        } else {
            // real samples: two consecutive outputs share a tap and read consecutive window samples (r + j, r + j + 1),
            // an aligned register pair when r + j is even.  Even taps accumulate into output pairs (0,1) .. (6,7), odd
            // taps into pairs (-1,0), (1,2) .. (7,8) (outer halves unused); the two sets are added once per step.
            static_assert(R == 8 && (C::LIVE % 2) == 0, "pairing below assumes 8 outputs per step");
            unsigned long long wp[C::LIVE / 2], aA[R / 2], aB[R / 2 + 1];
#pragma unroll
            for (int i = 0; i < C::LIVE / 2; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(wp[i]) : "f"(w1[2 * i]), "f"(w1[2 * i + 1]));
#pragma unroll
            for (int b = 0; b < R / 2; ++b) aA[b] = 0ull;
#pragma unroll
            for (int b = 0; b <= R / 2; ++b) aB[b] = 0ull;
#pragma unroll
            for (int j = 0; j < kDecTQ; ++j) {
                if (j % 2 == 0) {
#pragma unroll
                    for (int b = 0; b < R / 2; ++b) cfma(aA[b], t[j], wp[b + j / 2]);            // outputs (2b, 2b+1)
                } else {
#pragma unroll
                    for (int b = 0; b <= R / 2; ++b) cfma(aB[b], t[j], wp[b + (j - 1) / 2]);     // outputs (2b-1, 2b)
                }
            }
            float blo[R / 2 + 1], bhi[R / 2 + 1];
#pragma unroll
            for (int b = 0; b <= R / 2; ++b) asm("mov.b64 {%0, %1}, %2;" : "=f"(blo[b]), "=f"(bhi[b]) : "l"(aB[b]));
#pragma unroll
            for (int b = 0; b < R / 2; ++b) {
                float alo, ahi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(alo), "=f"(ahi) : "l"(aA[b]));
                acc[2 * b] = make_float2(alo + bhi[b], 0.f);
                acc[2 * b + 1] = make_float2(ahi + blo[b + 1], 0.f);
            }
        }
        // ---- reduce-scatter over the M residue lanes, one group of M outputs at a time: afterwards lane q holds
        // outputs sR + gM + q
#pragma unroll
        for (int g = 0; g < C::G; ++g) {
#pragma unroll
            for (int h = M / 2; h >= 1; h >>= 1) {
                const bool up = (q & h) != 0;
#pragma unroll
                for (int i = 0; i < h; ++i) {
                    const float2 keep = up ? acc[g * M + i + h] : acc[g * M + i];
                    const float2 send = up ? acc[g * M + i] : acc[g * M + i + h];
                    float2 got = make_float2(0.f, 0.f);
                    got.x = __shfl_xor_sync(0xffffffffu, send.x, h);
                    if constexpr (CPLX) got.y = __shfl_xor_sync(0xffffffffu, send.y, h);
                    acc[g * M + i] = make_float2(keep.x + got.x, keep.y + got.y);
                }
            }
            const long long k = k0 + (long long)s * R + g * M + q;
            const int c = ch0 + cl;
            if (k < P.N && c < nch) {
                if constexpr (CPLX) static_cast<float2 *>(yv)[(long long)c * ldy + k] = acc[g * M];
                else static_cast<float *>(yv)[(long long)c * ldy + k] = acc[g * M].x;
            }
        }

        // ---- this warp is done with group s (rows below (s+1)R): hand its slot back to the producer
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_base + 8 * (uint32_t)(s % NSLOT)) : "memory");
    }
    for (; g_waited <= glast; ++g_waited) {               // every issued load must have landed before exit
        mbar_wait(bar_base + 8 * w_slot, w_par);
        if (++w_slot == NSLOT) { w_slot = 0; w_par ^= 1u; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_decim8: M = 8, complex64 -- lane = channel, taps = launch constants
// ---------------------------------------------------------------------------------------------------------
// k_decim above spends a quarter of its FMA-pipe cycles and 40 % of its issue slots on everything that is not an FFMA2
// (window addressing, the shuffle reduce-scatter with its selects) and re-reads its window five times.  Here a LANE is
// one channel (16 channels x 2 output blocks per warp) and computes 8 consecutive outputs from a 320-sample window read
// ONCE with LDS.128 (two complex samples, [16 ch][16 samples] boxes with the 128-byte swizzle: conflict free), with
//     acc[r] += h[8 (m - r) + q] * x[8 m + q]        r = 0..7 outputs, window position 8 m + q
// unrolled over m and r: every tap sits at a fixed offset of the kernel's parameter block (constant bank; no tap table
// in registers or shared memory, no tap loads on the LSU), there is no cross-lane reduction, and an FFMA2 is 83 % of
// the instruction stream (13 per shared-memory load).  The price is shared memory: the outputs in flight need their 8 new
// samples each (64 B per output and channel), so a step of 64 outputs x 16 channels holds 64 KB of new samples; the ring
// is six 32 KB units (256 samples x 16 channels) with a full and an empty mbarrier each, filled by a producer warp that
// runs up to three units ahead.  The 2112 FFMA2 of a lane's step are split over a QUARTET of warps by q-pair (each warp
// reads its own quarter of the window); three helpers hand their partial sums to the first through shared memory (named
// barriers).  One persistent CTA per SM (16 consumer warps + producer) walks the work items round robin.
__device__ __forceinline__ void cfma_ordered(unsigned long long &acc, float t, unsigned long long x) {
    unsigned long long tt;
    asm("mov.b64 %0, {%1,%1};" : "=l"(tt) : "f"(t));
    asm volatile("fma.rn.f32x2 %0, %2, %1, %0;" : "+l"(acc) : "l"(tt), "l"(x));
}

constexpr int kD8Ch = 16;               // channels per CTA
constexpr int kD8Step = 64;             // outputs per channel and CTA step
constexpr int kD8TQ = 33;               // tap slots per residue (T <= 256, one slot of slack for the alignment shift)
constexpr int kD8Unit = 32768;          // ring unit: 256 samples x 16 channels x 8 B = 16 boxes of 2 KB
constexpr int kD8NU = 6;                // ring units: a step reads three, the producer runs up to three ahead
constexpr int kD8Scratch = 4 * 3 * 2048;   // partial sums of the three helper warps of every (a, b) quartet
constexpr int kD8Smem = kD8NU * kD8Unit + kD8Scratch + 128;

struct alignas(16) Dec8Params {
    long long k_begin, N;      // this launch covers outputs [k_begin, N)
    long long e;               // x index of the first sample of output 0's padded window (even)
    int KT;                    // outputs per work item (multiple of 64)
    int tiles, items, pad[3];  // items = channel groups x tiles; item i = (group i / tiles, tile i % tiles)
    float hq[4][kD8TQ + 1][2]; // hq[c][j][b] = padded hflip[8 j + 2 c + b]; rows of 34 slots at a 16-byte aligned offset (LDCU.128)
};

// One q-pair (window chunks 4 (m & 1) + C of every 16-sample box) of a lane's 8 outputs.  C is a RUN-TIME value: one copy of
// the 528-FFMA2 body (10 KB) serves the four q-pairs, the 66 taps are fetched from the parameter block with a register
// offset (LDC.64).  Measured alternatives: four compile-time copies (taps as LDCU / UR operands, 40 KB of code) miss the
// 32 KB instruction cache -- 27.4 Gout/s straight-line in every warp, 30.7 with one copy per warp quartet, against 39.8.
__device__ __forceinline__ void d8_body(const int C, unsigned long long (&acc)[8], const uint32_t (&base)[5], uint32_t sw, const Dec8Params &P) {
    const uint32_t xe = ((uint32_t)C ^ sw) << 4;                     // chunk C of the lane's row (m even)
    const uint32_t xo = ((uint32_t)(4 + C) ^ sw) << 4;               // chunk 4 + C (m odd)
    // volatile asm keeps the issue order as written: loads run PF window positions ahead of their FFMA2s, every FFMA2 is
    // followed by seven on other accumulators
    constexpr int PF = 4;
    unsigned long long xa[40], xb[40];
    auto lds = [&](int m) {
        // window position 8 m + 2 C (+1): group m / 8, box (m % 8) / 2, chunk 4 (m & 1) + C
        const uint32_t ad = base[m >> 3] + (uint32_t)(((m & 7) >> 1) * 2048) + ((m & 1) ? xo : xe);
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xa[m]), "=l"(xb[m]) : "r"(ad) : "memory");
    };
    // the taps of this q-pair, declared warp-uniform to the compiler (a shuffle from lane 0 of a value every lane holds):
    // FFMA2 then takes them from uniform registers -- with a vector-register scalar the FFMA2 reads five registers and
    // measured 70 % of its issue rate
    float t0[kD8TQ + 1], t1[kD8TQ + 1];
#pragma unroll
    for (int j = 0; j < kD8TQ + 1; j += 2) {                         // two slots = 16 bytes per fetch
        const float4 v = *reinterpret_cast<const float4 *>(&P.hq[C][j][0]);
        t0[j] = __shfl_sync(0xffffffffu, v.x, 0);
        t1[j] = __shfl_sync(0xffffffffu, v.y, 0);
        t0[j + 1] = __shfl_sync(0xffffffffu, v.z, 0);
        t1[j + 1] = __shfl_sync(0xffffffffu, v.w, 0);
    }
#pragma unroll
    for (int m = 0; m < PF; ++m) lds(m);
#pragma unroll
    for (int m = 0; m < 40; ++m) {
        if (m + PF < 40) lds(m + PF);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (m - r >= 0 && m - r < kD8TQ) cfma_ordered(acc[r], t0[m - r], xa[m]);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (m - r >= 0 && m - r < kD8TQ) cfma_ordered(acc[r], t1[m - r], xb[m]);
    }
}

// Persistent CTAs (one per SM) walk the work items round robin; the ring and its mbarrier phases run on across items (a
// global unit counter), so the producer is already fetching the next item while the consumers finish this one.
// Unit u of an item = samples [256 u, 256 u + 256) of its window; step s reads units 2s, 2s+1 (warps a = 0: window groups
// 0..7 of the step) or 2s+1, 2s+2 (warps a = 1: groups 4..11).  Every unit collects four "empty" arrivals: two from the
// a = 0 warps, two from the a = 1 warps; the item's first unit (not read by a = 1) and last unit (not read by a = 0) get
// the missing two as courtesy arrivals.
__global__ void __launch_bounds__(544, 1)
k_decim8(const __grid_constant__ CUtensorMap tmx, float2 *__restrict__ y, long long ldy, int nch,
         const __grid_constant__ Dec8Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler: the tap offsets become uniform
    const uint32_t ring = smem_u32(smem), scratch = ring + kD8NU * kD8Unit, full = scratch + kD8Scratch, empty = full + 8 * kD8NU;

    if (tid == 0) {
        if (ring & 1023u) __trap();
        for (int i = 0; i < kD8NU; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
    }
    __syncthreads();

    if (warp == 16) {
        if (lane == 0) {
            unsigned U = 0;                                          // units issued so far (all items)
            for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
                const int ch0 = (item / P.tiles) * kD8Ch;
                const long long k0 = P.k_begin + (long long)(item % P.tiles) * P.KT;
                const int ntile = (int)min((long long)P.KT, P.N - k0);
                const int nsteps = (ntile + kD8Step - 1) / kD8Step;
                const long long x0 = P.e + k0 * 8;                   // x index of the item's first sample (even, >= 0)
                for (int u = 0; u <= 2 * nsteps; ++u, ++U) {
                    const unsigned slot = U % kD8NU;
                    if (U >= kD8NU) mbar_wait(empty + 8 * slot, (U / kD8NU - 1) & 1u);
                    mbar_expect_tx(full + 8 * slot, kD8Unit);
                    for (int bx = 0; bx < 16; ++bx)
                        tma_load_2d(ring + slot * kD8Unit + (uint32_t)(bx * 2048), &tmx, (int)((x0 + u * 256 + bx * 16) * 2), ch0,
                                    full + 8 * slot);
                }
            }
        }
        return;
    }

    // ---- consumers: warp (a, b), lane half hb -> outputs 8 o .. 8 o + 7 of the step, o = 4 a + 2 hb + b; the lane's
    // window is the five 64-sample groups o .. o + 4 of the step (4 groups per unit)
    // Four warps share every (a, b): warp c of the quartet takes q-pair c (its own 40 of the 160 window loads) and the three
    // helpers hand their partial sums to the first through shared memory -- sixteen consumer warps, four per scheduler
    // (measured: 4 warps 37.9, 8 warps 39.6 Gout/s: the FMA pipe wants more than two warps to pick from).
    const int chl = lane & 15, hb = lane >> 4, c = warp & 3, pair = warp >> 2, a = pair >> 1;
    const int o = 4 * a + 2 * hb + (pair & 1);     // (c = warp & 3: a scheduler's four warps run ONE q-pair's code)
    unsigned gs = 0;                                                 // steps done so far (all items)
    const uint32_t lanerow = (uint32_t)(chl * 128), sw = (uint32_t)(chl & 7);
    unsigned U0 = 0;                                                 // unit counter at the start of the item
    for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
        const int c_glob = (item / P.tiles) * kD8Ch + chl;
        const long long k0 = P.k_begin + (long long)(item % P.tiles) * P.KT;
        const int ntile = (int)min((long long)P.KT, P.N - k0);
        const int nsteps = (ntile + kD8Step - 1) / kD8Step;
        float2 *yrow = y + (long long)c_glob * ldy;
        if (a == 1 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * (U0 % kD8NU)) : "memory");
        for (int s = 0; s < nsteps; ++s) {
            const unsigned ua = U0 + 2 * s + a;                      // first of the two units this warp reads
            mbar_wait(full + 8 * (ua % kD8NU), (ua / kD8NU) & 1u);
            mbar_wait(full + 8 * ((ua + 1) % kD8NU), ((ua + 1) / kD8NU) & 1u);
            uint32_t base[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int gg = o + i;
                base[i] = ring + ((U0 + 2 * s + (gg >> 2)) % kD8NU) * kD8Unit + (uint32_t)((gg & 3) * 8192) + lanerow;
            }
            unsigned long long acc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = 0ull;
            d8_body(c, acc, base, sw, P);
            // ---- this warp is done with its two units
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * (ua % kD8NU)) : "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * ((ua + 1) % kD8NU)) : "memory");
            }
            // ---- the quartet's partial sums meet: [pair][helper][r][lane]; named barriers 1 + 2 pair (data ready) and
            // 2 + 2 pair (scratch free again)
            const uint32_t sc = scratch + (uint32_t)(pair * 3 * 2048 + lane * 8);
            if (c != 0) {
                if (gs) asm volatile("bar.sync %0, 128;" ::"r"(2 + 2 * pair) : "memory");
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    asm volatile("st.shared.b64 [%0], %1;" ::"r"(sc + (uint32_t)((c - 1) * 2048 + r * 256)), "l"(acc[r]) : "memory");
                asm volatile("bar.arrive %0, 128;" ::"r"(1 + 2 * pair) : "memory");
                ++gs;
                continue;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + 2 * pair) : "memory");
#pragma unroll
            for (int w = 0; w < 3; ++w) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    unsigned long long v;
                    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(sc + (uint32_t)(w * 2048 + r * 256)) : "memory");
                    acc[r] = cadd(acc[r], v);
                }
            }
            asm volatile("bar.arrive %0, 128;" ::"r"(2 + 2 * pair) : "memory");
            ++gs;
            // ---- 8 consecutive outputs of one channel: 64 contiguous bytes
            const long long k = k0 + (long long)s * kD8Step + 8 * o;
            if (c_glob < nch) {
                if (k + 8 <= P.N && ((reinterpret_cast<uintptr_t>(yrow + k) & 15) == 0)) {
#pragma unroll
                    for (int r = 0; r < 8; r += 2) {
                        float4 v;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(acc[r]));
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(v.z), "=f"(v.w) : "l"(acc[r + 1]));
                        *reinterpret_cast<float4 *>(yrow + k + r) = v;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        float2 v;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(acc[r]));
                        if (k + r < P.N) yrow[k + r] = v;
                    }
                }
            }
        }
        if (a == 0 && lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * ((U0 + 2 * nsteps) % kD8NU)) : "memory");
        U0 += 2 * nsteps + 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct DecPlan {
    bool ok = false;
    bool cplx = true;                  // complex64 samples, else float32
    int M = 8;
    int64_t T = 0;
    DecParams *hp = nullptr;
    float *d_taps = nullptr;
    PFN_encodeTiled encode = nullptr;
    int num_sms = 148;
    // k_decim8 (M = 8, complex64): the padded taps for both alignments, copied into every launch's parameter block
    bool d8 = false;
    Dec8Params *hp8 = nullptr;
    float h8[2][4][kD8TQ + 1][2];
};

// hq tables of k_decim8 from the flipped taps (bank[i] multiplies window sample i)
static inline void decim8_set_bank(DecPlan &p, const std::vector<double> &bank) {
    if (!p.d8) return;
    memset(p.h8, 0, sizeof(p.h8));
    const int64_t Tp = (int64_t)kD8TQ * 8;
    for (int64_t delta = 0; delta < 2; ++delta) {
        const int64_t zf = Tp - p.T - delta;
        for (int64_t i = 0; i < p.T; ++i) {
            const int64_t ip = i + zf;
            p.h8[delta][(ip & 7) >> 1][ip >> 3][ip & 1] = (float)bank[(size_t)i];
        }
    }
}

static inline void decim_release(DecPlan &p) {
    delete p.hp;
    p.hp = nullptr;
    delete p.hp8;
    p.hp8 = nullptr;
    p.d8 = false;
    cudaFree(p.d_taps);
    p.d_taps = nullptr;
    p.ok = false;
}

// kind/tx/ty are the mrb.h enums (2 decimator ; 2 = complex64)
static inline int32_t decim_prepare(DecPlan &p, int kind, int tx, int ty, int64_t L, int64_t M, int64_t T,
                                    const std::vector<double> &bank, const cudaDeviceProp &prop) {
    p.ok = false;
    if (kind != 2 || tx != ty || !(tx == 2 || tx == 0) || L != 1) return 0;
    p.cplx = tx == 2;
    // float32: a TMA box row is M*4 bytes (>= 16) and the window start is aligned to 4 samples (M >= 4)
    if (!((p.cplx && M == 2) || M == 4 || M == 8) || T > (kDecTQ - 1) * M) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int32_t)(e ? e : cudaErrorUnknown);
    p.encode = (PFN_encodeTiled)fn;
    p.num_sms = prop.multiProcessorCount;
    p.M = (int)M; p.T = T;
    p.hp = new DecParams();
    memset(p.hp, 0, sizeof(DecParams));
    // padded taps: Tp = 33 M slots; zf zeros in front (they multiply samples older than the window), delta zeros
    // behind (they multiply samples newer than x[n_k]: finite data or TMA zero fill)
    const int64_t Tp = kDecTQ * M;
    const int64_t A = p.cplx ? 2 : 4;                                    // samples per 16 bytes
    std::vector<float> ht((size_t)A * kDecTQ * 32, 0.f);
    for (int64_t delta = 0; delta < A; ++delta) {
        const int64_t zf = Tp - T - delta;
        for (int64_t i = 0; i < T; ++i) {
            const int64_t ip = i + zf;
            ht[(size_t)((delta * kDecTQ + ip / M) * 32 + ip % M)] = (float)bank[i];   // bank = flipud(h): tap i multiplies window sample i
        }
    }
    e = cudaMalloc(&p.d_taps, ht.size() * sizeof(float));
    if (e != cudaSuccess) return (int32_t)e;
    e = cudaMemcpy(p.d_taps, ht.data(), ht.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int32_t)e;
    p.hp->taps = p.d_taps;
    if (p.cplx) {
        if (M == 4) e = cudaFuncSetAttribute(k_decim<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg<4, true>::SMEM);
        else if (M == 8) e = cudaFuncSetAttribute(k_decim<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg<8, true>::SMEM);
        else e = cudaFuncSetAttribute(k_decim<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg<2, true>::SMEM);
    } else {
        if (M == 4) e = cudaFuncSetAttribute(k_decim<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg<4, false>::SMEM);
        else e = cudaFuncSetAttribute(k_decim<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg<8, false>::SMEM);
    }
    if (e != cudaSuccess) return (int32_t)e;
    static const bool no8 = getenv("MRB_DECIM8") && atoi(getenv("MRB_DECIM8")) == 0;
    if (p.cplx && M == 8 && !no8 && kD8Smem <= (int)prop.sharedMemPerBlockOptin) {
        e = cudaFuncSetAttribute(k_decim8, cudaFuncAttributeMaxDynamicSharedMemorySize, kD8Smem);
        if (e != cudaSuccess) return (int32_t)e;
        p.hp8 = new Dec8Params();
        memset(p.hp8, 0, sizeof(Dec8Params));
        p.d8 = true;
        decim8_set_bank(p, bank);
    }
    p.ok = true;
    return 0;
}

// Live tap update on the device: the per-alignment residue tables from the flipped taps (bank[i] = hflip[i], float32),
// one thread per (alignment, tap).  Same layout as decim_prepare builds on the host.
__global__ void __launch_bounds__(256) k_decim_taps(const float *__restrict__ bank, int T, int M, int A, float *__restrict__ taps) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int total = A * kDecTQ * 32;
    if (idx >= total) return;
    // zero everything first: a slot (delta, j, q) holds padded tap ip = j*M + q, i.e. bank[ip - zf] when inside [0, T)
    const int delta = idx / (kDecTQ * 32), rem = idx - delta * kDecTQ * 32, j = rem / 32, q = rem & 31;
    float v = 0.f;
    if (q < M) {
        const int zf = kDecTQ * M - T - delta;
        const int i = j * M + q - zf;
        if (i >= 0 && i < T) v = bank[i];
    }
    taps[idx] = v;
}

static inline void decim_set_taps(DecPlan &p, const float *d_bank_f32, cudaStream_t st) {
    if (!p.ok) return;
    const int A = p.cplx ? 2 : 4;
    k_decim_taps<<<(A * kDecTQ * 32 + 255) / 256, 256, 0, st>>>(d_bank_f32, (int)p.T, p.M, A, p.d_taps);
}

// Launch for outputs [k_begin, N) of this chunk.  Returns k_begin (>= 0; the caller computes the outputs before it
// with the generic kernel), -1 when the call is not covered, -2 on a CUDA error.
static inline int64_t decim_try_launch(DecPlan &p, const GenParams &G, cudaStream_t st, const char **name,
                                       int64_t *launches) {
    static const bool trace = getenv("MRB_TRACE") != nullptr;
#define MRB_DEC_SKIP(why) do { if (trace) fprintf(stderr, "[mrb] decimator kernel not used: %s\n", why); return -1; } while (0)
    if (!p.ok) MRB_DEC_SKIP("configuration not covered");
    if (G.mode != SEQ_INTEGER || G.L != 1 || G.p0 != 0) MRB_DEC_SKIP("not a decimator schedule");
    const int64_t A = p.cplx ? 2 : 4;
    if (((uintptr_t)G.x & 15) || ((uintptr_t)G.y & (p.cplx ? 7 : 3)) || (G.ldx % A)) MRB_DEC_SKIP("alignment");
    if (G.n_in >= (1ll << 29) || G.nout >= (1ll << 29)) MRB_DEC_SKIP("size");
    const int M = p.M;
    const int64_t Tp = (int64_t)kDecTQ * M;
    // output k reads x[kM + d0m1 - (T-1) .. kM + d0m1]; with zf = Tp - T - delta zeros in front the padded window
    // starts at e = d0m1 - (T-1) - zf, and delta makes e a multiple of A samples (16-byte aligned TMA box starts)
    const int64_t e0 = G.d0m1 - (p.T - 1) - (Tp - p.T);
    const int64_t delta = ((-e0 % A) + A) % A;                  // e0 + delta is a multiple of A
    const int64_t e = e0 + delta;
    int64_t k_begin = e >= 0 ? 0 : ceil_div(-e, M);
    k_begin = (k_begin + 7) / 8 * 8;                           // whole steps: k_begin*M keeps e + k_begin*M aligned
    if (G.nout - k_begin < 64) MRB_DEC_SKIP("chunk too short");

    const int64_t span = G.nout - k_begin;
    if (p.d8 && ceil_div(G.nch, kD8Ch) * (span / kD8Step) >= 2 * p.num_sms && ((uintptr_t)G.y & 7) == 0) {
        // ---- lane-per-channel kernel with launch-constant taps
        Dec8Params &P8 = *p.hp8;
        P8.k_begin = k_begin; P8.N = G.nout; P8.e = e;
        memcpy(P8.hq, p.h8[delta], sizeof(P8.hq));
        const int64_t groups8 = ceil_div(G.nch, kD8Ch);
        // work items of `ts` steps, walked round robin by one persistent CTA per SM: pick the item length with the shortest
        // makespan (an item boundary costs a fraction of a step: the producer is already ahead)
        const int64_t total_steps = ceil_div(span, kD8Step);
        int64_t best_ts = 1;
        double best = 1e300;
        for (int64_t ts = 2; ts <= std::min<int64_t>(total_steps, 64); ++ts) {
            const int64_t items = groups8 * ceil_div(total_steps, ts);
            const double cost = (double)ceil_div(items, p.num_sms) * ((double)ts + 0.3);
            if (cost < best) { best = cost; best_ts = ts; }
        }
        P8.KT = (int)(best_ts * kD8Step);
        const int64_t tiles8 = ceil_div(span, P8.KT);
        P8.tiles = (int)tiles8; P8.items = (int)(tiles8 * groups8);
        CUtensorMap tm8;
        cuuint64_t dims[2] = {(cuuint64_t)(2 * G.n_in), (cuuint64_t)G.nch};
        cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 8};
        cuuint32_t box[2] = {32, (cuuint32_t)kD8Ch};
        cuuint32_t es[2] = {1, 1};
        if (p.encode(&tm8, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            MRB_DEC_SKIP("x tensor map");
        k_decim8<<<(unsigned)std::min<int64_t>(P8.items, p.num_sms), 544, kD8Smem, st>>>(tm8, static_cast<float2 *>(G.y), G.ldy, (int)G.nch, P8);
        if (cudaPeekAtLastError() != cudaSuccess) return -2;
        *name = "decim8_c64";
        ++*launches;
        return k_begin;
    }
    DecParams &P = *p.hp;
    P.k_begin = k_begin; P.N = G.nout; P.e = e; P.delta = (int)delta;
    const int64_t groups = ceil_div(G.nch, kDecRows);
    const int64_t R = 8;                                     // DecCfg<M>::R
    static const int wv = getenv("MRB_DEC_WAVES") ? atoi(getenv("MRB_DEC_WAVES")) : 4;
    int64_t tiles = std::max<int64_t>(1, std::min<int64_t>(span / (8 * R), ceil_div((int64_t)wv * 2 * p.num_sms, groups)));
    P.KT = (int)(ceil_div(ceil_div(span, tiles), R) * R);
    tiles = ceil_div(span, P.KT);

    CUtensorMap tmx;
    const int fpe = p.cplx ? 2 : 1;                            // floats per sample
    cuuint64_t dims[2] = {(cuuint64_t)(fpe * G.n_in), (cuuint64_t)G.nch};
    cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 4 * fpe};
    cuuint32_t box[2] = {(cuuint32_t)(fpe * M), kDecRows};
    cuuint32_t es[2] = {1, 1};
    if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        MRB_DEC_SKIP("x tensor map");
#undef MRB_DEC_SKIP
    dim3 grid((unsigned)tiles, (unsigned)groups);
    if (p.cplx) {
        if (M == 4) k_decim<4, true><<<grid, 32 * (DecCfg<4, true>::WARPS + 1), DecCfg<4, true>::SMEM, st>>>(tmx, G.y, G.ldy, (int)G.nch, P);
        else if (M == 8) k_decim<8, true><<<grid, 32 * (DecCfg<8, true>::WARPS + 1), DecCfg<8, true>::SMEM, st>>>(tmx, G.y, G.ldy, (int)G.nch, P);
        else k_decim<2, true><<<grid, 32 * (DecCfg<2, true>::WARPS + 1), DecCfg<2, true>::SMEM, st>>>(tmx, G.y, G.ldy, (int)G.nch, P);
    } else {
        if (M == 4) k_decim<4, false><<<grid, 32 * (DecCfg<4, false>::WARPS + 1), DecCfg<4, false>::SMEM, st>>>(tmx, G.y, G.ldy, (int)G.nch, P);
        else k_decim<8, false><<<grid, 32 * (DecCfg<8, false>::WARPS + 1), DecCfg<8, false>::SMEM, st>>>(tmx, G.y, G.ldy, (int)G.nch, P);
    }
    if (cudaPeekAtLastError() != cudaSuccess) return -2;
    *name = p.cplx ? (M == 4 ? "decim_c64_m4" : M == 8 ? "decim_c64_m8" : "decim_c64_m2") : (M == 4 ? "decim_f32_m4" : "decim_f32_m8");
    ++*launches;
    return k_begin;
}

}  // namespace mrb
