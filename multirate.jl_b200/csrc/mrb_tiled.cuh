// mrb_tiled.cuh -- the tiled fast path for integer-schedule kernels with unit input stride
// (FIRStandard, and FIRRational with L <= M < 2L such as 147//160 or M < L <= 1.5 M such as 160//147), complex64
// samples x float32 taps.
//
// What the hardware measurements forced (tools/ubench*.cu, profiles/README.md):
//  * lane = channel.  Every thread of a CTA walks the SAME outputs, so the phase / input-index bookkeeping of
//    src/Filters.jl:567-568 is warp-uniform and lives in the uniform datapath.
//  * taps: the flipped phase-major bank (taps2pfb, src/Filters.jl:284-298) rides in the __grid_constant__
//    parameter block; with a uniform row index ptxas emits LDCU c[0x0][UR+imm] and feeds FFMA2 a
//    uniform-register operand -- taps cost no shared-memory bandwidth and no vector registers.  One complex x
//    real FMA is one FFMA2.
//  * LDCU misses go to the GPC constant cache (GCC), which sustains only ~4 requests per clock CHIP-WIDE.  Every
//    CTA re-fetches each bank row once per L outputs, so the GCC traffic is (channel groups) x (outputs) x (row
//    bytes / 64) and falls only with the number of channels that share a fetch: a CTA is 128 channels, and its
//    warps consume the rows of the bank at the same time and strictly one row after the other (one basic block
//    per output; ptxas otherwise interleaves all rows of a run and thrashes the small per-SM constant cache).
//  * outputs are grouped into RUNS: maximal sets of <= RMAX consecutive outputs whose input index advances by
//    exactly one per output (the phase does not wrap inside a run).  Inside a run the window register index of
//    (output r, tap i) is the compile-time constant r+i, so the dot products (unsafedot, src/support.jl:5-14)
//    are straight-line FFMA2 on registers.  A run is split over two warps (outputs 0..5 / 6..11): 8 warps per
//    CTA, 2 CTAs per SM.
//  * samples: TMA (cp.async.bulk.tensor.2d, SWIZZLE_64B, 256-byte L2 promotion) streams [128 ch][8 samples]
//    boxes of x into a shared-memory ring; each thread reads its channel's window with conflict-free LDS.128.
//    The one CTA barrier of a run sits BEHIND the window loads: once every warp holds its window in registers,
//    all boxes before the next run's window are refilled -- a whole run ahead of their first use.
//  * results are staged in shared memory ([128 ch][8 outputs], SWIZZLE_64B, 4 buffers) and leave by TMA store.
//  * every shared-memory address is  base + (per-lane constant XOR uniform word);  the uniform words come from
//    two small tables in the parameter block, so a window load is LOP3 + LDS.128 and nothing else.
// Outputs whose window touches the history (the first ~T outputs of a chunk) and every configuration this
// kernel does not cover are computed by k_generic (mrb_kernels.cuh) -- still on the GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mrb_kernels.cuh"
#include "mrb_seq.h"

namespace mrb {

constexpr int kTiledRows = 128;         // channels per CTA
#ifndef MRB_BOX_SHIFT
#define MRB_BOX_SHIFT 4
#endif
constexpr int kBoxShift = MRB_BOX_SHIFT;                    // 3: 64-byte box rows (SWIZZLE_64B), 4: 128-byte (SWIZZLE_128B)
constexpr int kBoxSamples = 1 << kBoxShift;                 // samples per TMA box row
constexpr int kBoxBytes = kTiledRows * kBoxSamples * 8;     // 8192
constexpr int kOutChunk = 8;            // outputs per staged TMA store (64-byte rows)
constexpr int kOutBufs = 4;
constexpr int kOutBytes = kTiledRows * kOutChunk * 8;       // 8192
constexpr int kBankFloats = 6000;       // tap bank capacity in kernel-parameter space
constexpr int kMaxTiles = 192;          // time tiles per launch (their start states ride in parameter space)
constexpr int kMaxPhases = 1024;
constexpr int kTPAD = 24, kRMAX = 12, kNBOX = 80 / kBoxSamples;   // the ring holds 80 samples per channel
constexpr int kOB = 2;                  // outputs per basic block of the run body (1, 2 and 3 measure the same)
constexpr int kRingPairs = kBoxSamples / 2 * kNBOX;   // sample pairs (16-byte chunks per row) the ring holds
constexpr int kWinLen = (kRingPairs + (kTPAD + kRMAX) / 2 + 8 + 3) / 4 * 4;

struct alignas(16) TiledParams {
    long long k_begin, N;      // this launch covers outputs [k_begin, N)
    int L, M;
    int KT;                    // outputs per tile (multiple of 16)
    int pad0;
    // Start state of every time tile, computed on the host (closed form of src/Filters.jl:567-568) so that the
    // kernel's sequencing starts from parameter space and stays in the uniform datapath.
    //   j   : bank row of the tile's first output (rows are stored in RUN ORDER, see `bank`)
    //   s   : x-sample index of its window start, relative to box 0 of the tile
    //   xc0 : float coordinate of box 0 in the x tensor map
    struct Tile { int j, s, xc0, pad; } tile[kMaxTiles];
    // per bank row j: run length (bits 0-7, <= RMAX), "run ends on a phase wrap" (bit 8, M > L: the input index then
    // skips one extra sample; bit 9, M < L: it repeats one), row of the next run's first output (bits 16-31)
    int runtab[kMaxPhases];
    // win[c][i] = W((c + i) mod kRingPairs): W(u) = swizzle chunk bits | ring slot offset of sample pair u.  Four
    // copies shifted by c = 0..3 so that a window starting at any pair is fetched with aligned 128-bit LDCU.
    unsigned win[4][kWinLen];
    // wout[i] = address word of staged output (i mod 32): chunk | odd/even | buffer
    unsigned wout[48];
    // Tap bank in RUN ORDER: row j holds branch phi_j = (j * (M-L)) mod L of the flipped phase-major bank
    // (taps2pfb, src/Filters.jl:284-298), left-padded with zeros to TPAD taps; consecutive outputs of a run read
    // consecutive rows, so every tap address inside a run is (one uniform base) + (compile-time offset).
    // Rows L .. L+RMAX-2 repeat rows 0 .. RMAX-2.
    float4 bank[kBankFloats / 4];
};
static_assert(sizeof(TiledParams) + 2 * sizeof(CUtensorMap) <= 32764, "kernel parameter space is 32764 bytes");

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// complex x real FMA: acc(re,im) += t * x(re,im)  -> one FFMA2 with a scalar (uniform-register) tap operand
__device__ __forceinline__ void cfma(unsigned long long &acc, float t, unsigned long long x) {
    unsigned long long tt;
    asm("mov.b64 %0, {%1,%1};" : "=l"(tt) : "f"(t));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(tt), "l"(x));
}
__device__ __forceinline__ unsigned long long cadd(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// this warp's window of the run -> registers: NP aligned sample pairs, conflict-free LDS.128 under SWIZZLE_64B
// ---------------------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ void load_window(unsigned long long (&xw)[2 * NP], const unsigned *wt, uint32_t in_base,
                                            uint32_t rowpart) {
#pragma unroll
    for (int q = 0; q < NP; q += 4) {
        const uint4 w4 = *reinterpret_cast<const uint4 *>(wt + q);
        const unsigned w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (q + e < NP) {
                const uint32_t a = in_base + (rowpart ^ w[e]);
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];"
                             : "=l"(xw[2 * (q + e)]), "=l"(xw[2 * (q + e) + 1]) : "r"(a) : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// RW outputs of one run (the first `len` are kept) from the register window; j / kpos = bank row / tile-relative index
// of the first output this warp computes, DELTA = parity of the run's window start.
// ---------------------------------------------------------------------------------------------------------
template <int TPAD, int RW, int DELTA, int OB>
__device__ __forceinline__ void run_body(const TiledParams &P, const unsigned long long (&xw)[(TPAD + RW + 1) / 2 * 2],
                                         uint32_t rowpart, int j, int len, int kpos, uint32_t out_base) {
    static_assert(RW % OB == 0, "whole blocks");
    const float4 *rows4 = P.bank + j * (TPAD / 4);         // uniform base; everything below is base + constant
    const int o0 = kpos & 31;
#pragma unroll
    for (int rb = 0; rb < RW; rb += OB) {
        // One basic block per OB outputs (the uniform early exit below ends it): the rows of the bank are consumed
        // in order, at most OB at a time -- ptxas otherwise interleaves all rows of the run and thrashes the small
        // per-SM constant cache -- and a short run does not pay for the outputs it does not have.
#pragma unroll
        for (int r = rb; r < rb + OB; ++r) {
            unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
#pragma unroll
            for (int i = 0; i < TPAD; i += 4) {
                const float4 t = rows4[r * (TPAD / 4) + i / 4];
                cfma(a0, t.x, xw[DELTA + r + i]);
                cfma(a1, t.y, xw[DELTA + r + i + 1]);
                cfma(a2, t.z, xw[DELTA + r + i + 2]);
                cfma(a3, t.w, xw[DELTA + r + i + 3]);
            }
            const unsigned long long y = cadd(cadd(a0, a1), cadd(a2, a3));
            if (OB == 1 || r < len) {
                const uint32_t a = out_base + (rowpart ^ P.wout[o0 + r]);
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(y) : "memory");
            }
        }
        if (rb + OB >= len) break;
    }
}

// ---------------------------------------------------------------------------------------------------------
// kernel.  grid = (time tiles, channel groups of 128), block = 256 threads: warp w computes half (w >> 2) of every
// run for channels 32 (w & 3) .. 32 (w & 3) + 31 of the group.
// ---------------------------------------------------------------------------------------------------------
template <int TPAD, int RMAX, int NBOX, int OB>
__global__ void __launch_bounds__(256, 2)
k_tiled_c64(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
            const __grid_constant__ TiledParams P) {
    static_assert(TPAD % 4 == 0 && RMAX % 4 == 0, "taps are fetched four at a time; both halves start on even outputs");
    static_assert(7 + 2 * RMAX <= kOutChunk * kOutBufs, "staging buffers: see the store hazard note in the loop");
    constexpr int RW = RMAX / 2;                          // outputs per warp per run
    constexpr int NP = (TPAD + RW + 1) / 2;               // sample pairs one warp's outputs touch
    constexpr int NPRUN = (TPAD + RMAX + 1) / 2;          // sample pairs the whole run touches
    static_assert(2 * NPRUN + kBoxSamples - 1 <= (NBOX - 1) * kBoxSamples, "the ring must hold a run's window and leave a box to refill");
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *in_ring = smem;                                   // NBOX boxes [128][8] complex64, SWIZZLE_64B
    unsigned char *out_ring = smem + NBOX * kBoxBytes;               // kOutBufs chunks [128][8] complex64, SWIZZLE_64B
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(out_ring + kOutBufs * kOutBytes);

    const int tid = threadIdx.x;
    const int row = tid & (kTiledRows - 1);                          // channel within the group
    const int half = __shfl_sync(0xffffffffu, tid >> 7, 0);          // warp-uniform by construction (keeps taps on LDCU)
    const int ch0 = blockIdx.y * kTiledRows;
    const uint32_t in_base = smem_u32(in_ring), out_base = smem_u32(out_ring), bar_base = smem_u32(bars);
    // SWIZZLE_64B: the 16-byte chunk index is XORed with (row >> 1) & 3; SWIZZLE_128B: with row & 7.  The per-lane
    // part of every address, for the output staging buffers (64-byte rows) and for the input ring:
    const uint32_t rowpart = ((uint32_t)row * 64u) ^ ((((uint32_t)row >> 1) & 3u) << 4);
    const uint32_t rowpart_in = kBoxShift == 3 ? rowpart : ((uint32_t)row * 128u) ^ (((uint32_t)row & 7u) << 4);

    // ---- tile start state: from parameter space (uniform)
    const int ka_rel = blockIdx.x * P.KT;                              // relative to k_begin
    const int ntile = min(P.KT, (int)(P.N - P.k_begin) - ka_rel);
    int j = P.tile[blockIdx.x].j;                                      // bank row (run order) of the next output
    int s = P.tile[blockIdx.x].s;                                      // window start, relative to box 0 of the tile
    const int xc0 = P.tile[blockIdx.x].xc0;                            // float coordinate of box 0 in tmx
    const int yc0 = ((int)P.k_begin + ka_rel) * 2;
    // boxes this tile is expected to touch (prefetch bound); demand may exceed it by a box or two
    const int jend = ((ntile + (int)(((long long)ntile * (P.M - P.L)) / P.L) + TPAD + RMAX) >> kBoxShift) + 1;

    if (tid == 0) {
        if (in_base & 1023u) __trap();                               // the swizzle formulas assume 1 KiB alignment
#pragma unroll 1
        for (int i = 0; i < NBOX; ++i) mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
#pragma unroll 1
        for (int jj = 0; jj < NBOX; ++jj) {                          // prologue: fill the ring
            mbar_expect_tx(bar_base + 8 * jj, kBoxBytes);
            tma_load_2d(in_base + (uint32_t)(jj * kBoxBytes), &tmx, xc0 + jj * (2 * kBoxSamples), ch0, bar_base + 8 * jj);
        }
    }
    __syncthreads();

    int k = 0;                           // tile-relative index of the next output
    int j_issued = NBOX, i_slot = 0;     // boxes issued so far (tile-relative) and the ring slot of the next one
    int j_waited = 0, w_slot = 0;        // boxes already waited for, ring slot of the next one
    uint32_t w_par = 0;                  //   ... and its mbarrier phase parity
    int q_flushed = 0;                   // output chunks already handed to TMA

    while (k < ntile) {
        const int rt = P.runtab[j];
        const int len = min(rt & 0xff, ntile - k);
        const int A = s & ~1;                                  // aligned window start
        const int jneed = (A + 2 * NPRUN - 1) >> kBoxShift;    // newest box the run's windows touch
        const int q_done = k >> 3;                             // chunks completed by earlier runs
        const int p = ((A >> 1) + half * (RW / 2)) % kRingPairs;   // ring position (in pairs) of this warp's window
        const unsigned *wt = P.win[p & 3] + (p & ~3);

        // ---- this warp's window -> registers
#pragma unroll 1
        for (; j_waited <= jneed; ++j_waited) {
            mbar_wait(bar_base + 8 * w_slot, w_par);
            if (++w_slot == NBOX) { w_slot = 0; w_par ^= 1u; }
        }
        unsigned long long xw[2 * NP];
        load_window<NP>(xw, wt, in_base, rowpart_in);

        // ---- the one barrier of the run.  Behind it (a) every warp holds its window in registers, so all boxes
        // before the NEXT run's window are free and are refilled now, a whole run ahead of their first use; (b) every
        // warp has staged the previous run's outputs, so completed chunks can leave.
        // the run ended on a phase wrap: the input index then skips one sample (M > L) or repeats one (M < L)
        const int s_next = s + len + ((rt >> 8) & 1) - ((rt >> 9) & 1);
        const bool has_next = k + len < ntile;
        const int jA_next = (has_next ? s_next : s) >> kBoxShift;
        // the next run's windows must be on their way after this barrier whatever the estimate `jend` says
        const int jneed_next = has_next ? ((s_next & ~1) + 2 * NPRUN - 1) >> kBoxShift : jneed;
        const bool flush = q_done > q_flushed;
        if (flush) fence_async_smem();                         // this thread's st.shared -> visible to the async proxy
        if (tid == 0) tma_wait_read<0>();                      // stores issued a run ago have left their buffers
        __syncthreads();
        const int jtarget = min(max(jneed_next, min(jA_next + NBOX - 1, jend)), jA_next + NBOX - 1);
        if (tid == 0) {
            int sl = i_slot;
#pragma unroll 1
            for (int jj = j_issued; jj <= jtarget; ++jj) {
                const uint32_t bar = bar_base + 8 * sl;
                mbar_expect_tx(bar, kBoxBytes);
                tma_load_2d(in_base + (uint32_t)(sl * kBoxBytes), &tmx, xc0 + jj * (2 * kBoxSamples), ch0, bar);
                if (++sl == NBOX) sl = 0;
            }
            // The run writes chunks q_done.. ; the stores issued here read the buffers of chunks q_flushed..q_done-1
            // (q_flushed = chunk of the previous run's first output).  Two runs span <= 7 + 2*RMAX = 31 outputs from
            // the start of chunk q_flushed, i.e. at most 4 chunks: with 4 buffers the two sets never share a
            // buffer, and everything older was confirmed by the wait above.
            if (flush) {
#pragma unroll 1
                for (int q = q_flushed; q < q_done; ++q) {
                    tma_store_2d(&tmy, yc0 + q * 16, ch0, out_base + (uint32_t)((q & (kOutBufs - 1)) * kOutBytes));
                    tma_commit();
                }
            }
        }
        if (jtarget >= j_issued) {
            i_slot = (i_slot + (jtarget + 1 - j_issued)) % NBOX;
            j_issued = jtarget + 1;
        }
        q_flushed = q_done;

        // both halves run the same code (their windows were loaded RW samples apart): first row / output / count of
        // this warp's part of the run are uniform values, not template parameters -- half the instruction footprint
        if (len > half * RW) {
            const int r0 = half * RW;
            if (s & 1) run_body<TPAD, RW, 1, OB>(P, xw, rowpart, j + r0, len - r0, k + r0, out_base);
            else run_body<TPAD, RW, 0, OB>(P, xw, rowpart, j + r0, len - r0, k + r0, out_base);
        }

        // ---- advance the (uniform) schedule by `len` outputs
        k += len;
        s = s_next;
        j = rt >> 16;
    }

    // ---- drain: every issued load must have landed before the CTA gives its shared memory back
#pragma unroll 1
    for (; j_waited < j_issued; ++j_waited) {
        mbar_wait(bar_base + 8 * w_slot, w_par);
        if (++w_slot == NBOX) { w_slot = 0; w_par ^= 1u; }
    }
    // ---- flush the remaining chunks (the last one may be partial: TMA clips at the tensor bound N)
    fence_async_smem();
    if (tid == 0) tma_wait_read<0>();
    __syncthreads();
    if (tid == 0) {
        const int q_end = (ntile + kOutChunk - 1) >> 3;
#pragma unroll 1
        for (int q = q_flushed; q < q_end; ++q) {
            tma_store_2d(&tmy, yc0 + q * 16, ch0, out_base + (uint32_t)((q & (kOutBufs - 1)) * kOutBytes));
            tma_commit();
        }
        tma_wait_read<0>();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TiledPlan {
    bool ok = false;               // configuration is covered by a tiled kernel
    int tpad = 0, rmax = 0;
    TiledParams *hp = nullptr;     // host template of the parameter block (bank, run table, address tables filled once)
    PFN_encodeTiled encode = nullptr;
    int T = 0;
    int kt_min = 1024;             // MRB_TILED_KT forces the time tile (tuning)
    bool kt_forced = false;
    int num_sms = 148;
    std::vector<int> row_of_phase;
};

constexpr int kTiledThreads = 256;
constexpr int kTiledSmem = kNBOX * kBoxBytes + kOutBufs * kOutBytes + 8 * kNBOX;

static inline void tiled_release(TiledPlan &p) {
    delete p.hp;
    p.hp = nullptr;
    p.ok = false;
}

// kind/tx/ty are the mrb.h enums (0 standard, 3 rational ; 2 = complex64)
static inline int32_t tiled_prepare(TiledPlan &p, int kind, int tx, int ty, int64_t L, int64_t M, int64_t Nphi,
                                    int64_t T, const std::vector<double> &bank, const std::vector<double> &,
                                    const cudaDeviceProp &prop) {
    p.ok = false;
    p.num_sms = prop.multiProcessorCount;
    const bool kind_ok = kind == 0 /*standard*/ || kind == 3 /*rational*/;
    if (!kind_ok || tx != 2 || ty != 2) return 0;
    // unit input stride between consecutive outputs except at phase wraps: M in [L, 2L), or M < L with runs that are
    // not too short on average (L / (L - M) >= 3 outputs)
    const bool ratio_ok = (L <= M && M < 2 * L) || (M < L && 3 * (L - M) <= L);
    if (!ratio_ok || L > kMaxPhases || T > kTPAD || (L + kRMAX) * kTPAD > kBankFloats) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int32_t)(e ? e : cudaErrorUnknown);
    p.encode = (PFN_encodeTiled)fn;
    p.tpad = kTPAD; p.rmax = kRMAX; p.T = (int)T;
    p.hp = new TiledParams();
    memset(p.hp, 0, sizeof(TiledParams));
    p.hp->L = (int)L; p.hp->M = (int)M;
    if (const char *ev = getenv("MRB_TILED_KT")) { p.kt_min = std::max(64, atoi(ev) / 16 * 16); p.kt_forced = true; }
    // shared-memory address words (see load_window / run_body)
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < kWinLen; ++i) {
            const unsigned u = (unsigned)(c + i) % (unsigned)kRingPairs;
            p.hp->win[c][i] = ((u & (unsigned)(kBoxSamples / 2 - 1)) << 4) | ((u >> (kBoxShift - 1)) * (unsigned)kBoxBytes);
        }
    for (int i = 0; i < 48; ++i) {
        const unsigned kk = (unsigned)i & 31u;
        p.hp->wout[i] = (((kk >> 1) & 3u) << 4) | ((kk & 1u) << 3) | (((kk >> 3) & 3u) * (unsigned)kOutBytes);
    }
    const int64_t mp = M - L;                                          // phase step per output (mod L); < 0 when M < L
    p.row_of_phase.assign((size_t)L, 0);
    std::vector<int64_t> phase_of_row((size_t)L);
    for (int64_t j = 0; j < L; ++j) {
        phase_of_row[j] = (j * M) % L;                                 // a bijection: gcd(M, L) == 1
        p.row_of_phase[phase_of_row[j]] = (int)j;
    }
    float *hb = reinterpret_cast<float *>(p.hp->bank);
    for (int64_t j = 0; j < L + kRMAX - 1; ++j) {
        const int64_t ph = phase_of_row[j % L];
        // row, left-padded with zeros: padded tap i multiplies the sample (TPAD-1-i) before the window's last one
        for (int64_t i = 0; i < T; ++i) hb[j * kTPAD + (kTPAD - T) + i] = (float)bank[ph * T + i];
    }
    for (int64_t j = 0; j < L; ++j) {
        const int64_t ph = phase_of_row[j];
        // M >= L: outputs at phases ph, ph+mp, ... read consecutive input windows until the phase wraps past L (the
        // input index then advances by two).  M < L: phases ph, ph-|mp|, ... until the phase drops below |mp| (the
        // next output then reads the same window again).
        const int64_t to_wrap = mp == 0 ? kRMAX : mp > 0 ? (L - 1 - ph) / mp + 1 : ph / (-mp) + 1;
        const int64_t len = std::min<int64_t>(to_wrap, kRMAX);
        const int64_t wrap = (mp != 0 && len == to_wrap) ? (mp > 0 ? 1 : 2) : 0;
        p.hp->runtab[j] = (int)(len | (wrap << 8) | (((j + len) % L) << 16));
    }
    e = cudaFuncSetAttribute(k_tiled_c64<kTPAD, kRMAX, kNBOX, kOB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiledSmem);
    if (e != cudaSuccess) return (int32_t)e;
    p.ok = true;
    return 0;
}

// Live tap update: rewrite the run-ordered bank of an existing plan (host parameter block; the next launch carries it).
static inline void tiled_set_bank(TiledPlan &p, int64_t L, int64_t T, const std::vector<double> &bank) {
    if (!p.ok) return;
    float *hb = reinterpret_cast<float *>(p.hp->bank);
    for (int64_t j = 0; j < L + kRMAX - 1; ++j) {
        const int64_t ph = ((j % L) * (int64_t)p.hp->M) % L;            // the branch run-order row j holds (tiled_prepare)
        for (int64_t i = 0; i < T; ++i) hb[j * kTPAD + (kTPAD - T) + i] = (float)bank[ph * T + i];
    }
}

// Launch the tiled kernel for outputs [k_begin, N) of this chunk.  Returns k_begin (>= 0; the caller computes
// [0, k_begin) with the generic kernel), -1 when the call is not covered, -2 on a CUDA error.
static inline int64_t tiled_try_launch(TiledPlan &p, const GenParams &G, cudaStream_t st, const char **name,
                                       int64_t *launches) {
    if (!p.ok || G.mode != SEQ_INTEGER) return -1;
    // TMA needs 16-byte aligned bases and row pitches, and 32-bit coordinates
    if (((uintptr_t)G.x & 15) || ((uintptr_t)G.y & 15) || (G.ldx & 1) || (G.ldy & 1)) return -1;
    if (G.n_in >= (1ll << 30) || G.nout >= (1ll << 30)) return -1;
    // first output whose (padded) window lies entirely inside x: d0m1 + floor((p0 + k M)/L) >= TPAD-1
    const int64_t q = (int64_t)p.tpad - 1 - G.d0m1;
    int64_t kstar = q <= 0 ? 0 : ceil_div(q * G.L - G.p0, G.M);
    if (kstar < 0) kstar = 0;
    const int64_t k_begin = (kstar + 15) / 16 * 16;
    if (G.nout - k_begin < 64) return -1;                     // too small to be worth a tiled launch

    TiledParams &P = *p.hp;
    P.k_begin = k_begin; P.N = G.nout;
    const int64_t span = G.nout - k_begin;
    // Time tiles: ~1024..2048 outputs each, their number chosen so that the grid is a whole number of waves of
    // (2 CTAs per SM) wherever the shape allows (a 0.76-wave tail costs 2 % at BASELINE configs[4]).
    const int64_t groups = ceil_div(G.nch, kTiledRows), resident = 2 * (int64_t)p.num_sms;
    int64_t best_t = ceil_div(span, std::max(p.kt_min, 16));
    if (!p.kt_forced) {
        double best_eff = -1.0;
        const int64_t t_lo = std::max<int64_t>(1, ceil_div(span, 2048)), t_hi = std::max<int64_t>(t_lo, span / 768);
        for (int64_t t = t_lo; t <= std::min<int64_t>(t_hi, kMaxTiles); ++t) {
            const int64_t ctas = t * groups;
            const double eff = (double)ctas / (double)(ceil_div(ctas, resident) * resident);
            if (eff > best_eff + 1e-9) { best_eff = eff; best_t = t; }
        }
    }
    best_t = std::min<int64_t>(best_t, kMaxTiles);
    P.KT = (int)((ceil_div(span, best_t) + 15) / 16 * 16);
    const int64_t ntiles = ceil_div(span, P.KT);
    for (int64_t i = 0; i < ntiles; ++i) {
        const int64_t ka = k_begin + i * P.KT;
        const int64_t t0 = G.p0 + ka * G.M;
        const int64_t xs0 = G.d0m1 + t0 / G.L - (p.tpad - 1);         // x-sample index of the first window start
        const int64_t box0 = xs0 >> kBoxShift;
        P.tile[i].j = p.row_of_phase[(size_t)(t0 % G.L)];
        P.tile[i].s = (int)(xs0 - (box0 << kBoxShift));
        P.tile[i].xc0 = (int)(box0 << kBoxShift) * 2;
    }
    CUtensorMap tmx, tmy;
    {
        cuuint64_t dims[2] = {(cuuint64_t)(2 * G.n_in), (cuuint64_t)G.nch};
        cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 8};
        cuuint32_t box[2] = {2 * kBoxSamples, kTiledRows};
        cuuint32_t es[2] = {1, 1};
        // 256-byte L2 promotion: DRAM sees 256-byte bursts per channel row although a box row is 64 bytes
        // (tools/ubench3.cu: +10 % read bandwidth for this access pattern)
        if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, kBoxShift == 3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        cuuint64_t ydims[2] = {(cuuint64_t)(2 * G.nout), (cuuint64_t)G.nch};
        cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * 8};
        cuuint32_t ybox[2] = {2 * kOutChunk, kTiledRows};
        if (p.encode(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G.y, ydims, ystrides, ybox, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
    }
    dim3 grid((unsigned)ntiles, (unsigned)ceil_div(G.nch, kTiledRows));
    k_tiled_c64<kTPAD, kRMAX, kNBOX, kOB><<<grid, kTiledThreads, kTiledSmem, st>>>(tmx, tmy, P);
    if (cudaPeekAtLastError() != cudaSuccess) return -2;
    *name = "tiled_c64_t24_r12";
    ++*launches;
    return k_begin;
}

}  // namespace mrb
