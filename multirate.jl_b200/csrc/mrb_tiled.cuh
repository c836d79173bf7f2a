// mrb_tiled.cuh -- the tiled fast path for integer-schedule kernels with unit input stride
// (FIRStandard, and FIRRational with L <= M < 2L such as 147//160), complex64 or float32 samples.
//
// Mapping (B200, sm_100a), measured pipe rates in tools/ubench*.cu and DESIGN.md:
//  * lane = channel.  A CTA is 64 channels x 2 run halves = 4 warps; every thread walks the SAME runs, so the whole
//    phase / input-index bookkeeping (src/Filters.jl:567-568) is warp-uniform and lives in the uniform
//    datapath (UIADD3 / UISETP), never in vector registers.
//  * taps: the flipped phase-major bank (taps2pfb, src/Filters.jl:284-298) is a __grid_constant__ kernel
//    parameter; with a uniform phase index ptxas emits LDCU.64 c[0x0][UR+imm] and feeds FFMA2 a
//    uniform-register operand -- taps cost no shared-memory bandwidth and no vector registers.
//  * samples: TMA (cp.async.bulk.tensor.2d, SWIZZLE_64B) streams [64 channels][8 samples] boxes of x into a
//    shared-memory ring; each thread reads its channel's window with conflict-free LDS.128 into registers.
//  * outputs are grouped into RUNS: maximal sets of <= RMAX consecutive outputs whose input index advances
//    by exactly one per output (the phase does not wrap inside a run).  Inside a run the window register
//    index of (output r, tap i) is the compile-time constant r+i, so the dot products (unsafedot,
//    src/support.jl:5-14) are straight-line FFMA2 on registers: one complex x real FMA = one FFMA2.
//  * results are staged in shared memory ([64 channels][8 outputs], SWIZZLE_64B) and written with TMA stores,
//    so global writes are full 64-byte rows instead of 8-byte scatters.
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

constexpr int kTiledRows = 64;          // channels per CTA
constexpr int kTiledThreads = 128;      // 4 warps: (channel half) x (run half)
constexpr int kBoxSamples = 8;          // samples per TMA box row (64 B for complex64)
constexpr int kBankFloats = 6144;       // tap bank capacity in kernel-parameter space (24 KiB)
constexpr int kMaxTiles = 192;          // time tiles per launch (their start states ride in parameter space)
constexpr int kMaxPhases = 1024;
constexpr int kOutBufs = 8;             // staging buffers of kOutChunk outputs each
constexpr int kOutChunk = 4;            // outputs per staged TMA store (32-byte rows, SWIZZLE_32B)

struct TiledParams {
    long long k_begin, N;      // this launch covers outputs [k_begin, N)
    int L, M;
    int KT;                    // outputs per tile (multiple of 16)
    int pf_dist;               // L2 prefetch distance in 8-box groups (0 = off)
    // Start state of every time tile, computed on the host (closed form of src/Filters.jl:567-568) so that the
    // kernel's sequencing starts from parameter space and stays in the uniform datapath.
    //   j   : bank row of the tile's first output (rows are stored in RUN ORDER, see `bank`)
    //   s   : x-sample index of its window start, relative to box 0 of the tile
    //   xc0 : float coordinate of box 0 in the x tensor map
    struct Tile { int j, s, xc0, pad; } tile[kMaxTiles];
    // per bank row j: run length (bits 0-7, <= RMAX), "run ends on a phase wrap" (bit 8: the input index then
    // skips one extra sample), row of the next run's first output (bits 16-31)
    int runtab[kMaxPhases];
    // Tap bank in RUN ORDER: row j holds branch phi_j = (j * (M-L)) mod L of the flipped phase-major bank
    // (taps2pfb, src/Filters.jl:284-298), left-padded with zeros to TPAD taps; consecutive outputs of a run read
    // consecutive rows, so every tap address inside a run is (one uniform base) + (compile-time offset).
    // Rows L .. L+RMAX-2 repeat rows 0 .. RMAX-2.
    float bank[kBankFloats];
};

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
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
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
// one run: RMAX outputs (the first `len` are kept) from a register window.  DELTA = parity of the window start.
// ---------------------------------------------------------------------------------------------------------
// RW = outputs per warp per run; R0 = first output of the run this warp computes (the run is split over the
// two warps that share a channel group, which halves the shared-memory footprint per warp).
template <int TPAD, int RW, int R0, int DELTA, int NBOX>
__device__ __forceinline__ void run_body_c64(const TiledParams &P, int A, uint32_t ibase_x, int j, int len, int kpos,
                                             uint32_t obase_x) {
    // ---- register window: NP aligned sample pairs (LDS.128, conflict free under SWIZZLE_64B).  Every index
    // below is uniform, so the only per-thread work per load is one XOR with the folded base.
    constexpr int NP = (TPAD + RW + 1 + 1) / 2;
    unsigned long long xw[2 * NP];
    {
        const int u0 = (A >> 1) + R0 / 2;                 // pair index of this warp's window; 4 pairs per box
#pragma unroll
        for (int jj = 0; jj < NP; ++jj) {
            const int u = u0 + jj;
            const uint32_t a = (ibase_x ^ (uint32_t)((u & 3) << 4)) + (uint32_t)(((u >> 2) & (NBOX - 1)) << 12);
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xw[2 * jj]), "=l"(xw[2 * jj + 1]) : "r"(a));
        }
    }
    const float *rows = P.bank + (j + R0) * TPAD;         // uniform base; everything below is base + constant
    len -= R0;
    kpos += R0;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
#pragma unroll
        for (int i = 0; i < TPAD; i += 4) {
            const float2 t0 = *reinterpret_cast<const float2 *>(rows + r * TPAD + i);
            const float2 t1 = *reinterpret_cast<const float2 *>(rows + r * TPAD + i + 2);
            cfma(a0, t0.x, xw[DELTA + r + i]);
            cfma(a1, t0.y, xw[DELTA + r + i + 1]);
            cfma(a2, t1.x, xw[DELTA + r + i + 2]);
            cfma(a3, t1.y, xw[DELTA + r + i + 3]);
        }
        const unsigned long long y = cadd(cadd(a0, a1), cadd(a2, a3));
        if (r < len) {                                    // uniform predicate
            const int kk = kpos + r;                      // tile-relative output index
            const uint32_t a = (obase_x ^ (uint32_t)(((kk >> 1) & 1) << 4)) +
                               (uint32_t)((((kk >> 2) & (kOutBufs - 1)) << 11) + ((kk & 1) << 3));
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(y) : "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// kernel: complex64 samples, float32 taps.  grid = (time tiles, channel groups of 64), block = 64 threads.
// ---------------------------------------------------------------------------------------------------------
template <int TPAD, int RMAX, int NBOX>
__global__ void __launch_bounds__(kTiledThreads)
k_tiled_c64(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
            const __grid_constant__ CUtensorMap tmp, const __grid_constant__ TiledParams P) {
    static_assert((NBOX & (NBOX - 1)) == 0, "ring size must be a power of two");
    static_assert(RMAX % 4 == 0, "the run is split in two even halves");
    constexpr int RW = RMAX / 2;                          // outputs per warp per run
    constexpr int NPRUN = (TPAD + RMAX + 1 + 1) / 2;      // sample pairs the whole run touches
    constexpr int BOX_BYTES = kTiledRows * kBoxSamples * 8;   // 4096
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *in_ring = smem;                                   // NBOX boxes [64][8] complex64, SWIZZLE_64B
    unsigned char *out_ring = smem + NBOX * BOX_BYTES;               // kOutBufs chunks [64][4] complex64, SWIZZLE_32B
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(out_ring + kOutBufs * 2048);

    const int tid = threadIdx.x;
    const int ch0 = blockIdx.y * kTiledRows;
    const uint32_t in_base = smem_u32(in_ring), out_base = smem_u32(out_ring), bar_base = smem_u32(bars);
    const int row = tid & (kTiledRows - 1);                          // channel within the group
    const int half = tid >> 6;                                       // which half of every run this warp computes
    // SWIZZLE_64B: the 16-byte chunk index is XORed with (row>>1)&3.  Fold the per-thread part into the bases,
    // so that every shared-memory address is (per-thread base ^ uniform chunk bits) + uniform offset.
    const uint32_t row_swz4 = (((uint32_t)row >> 1) & 3u) << 4;
    const uint32_t ibase_x = (in_base + (uint32_t)row * 64u) ^ row_swz4;
    const uint32_t obase_x = (out_base + (uint32_t)row * 32u) ^ ((((uint32_t)row >> 2) & 1u) << 4);   // SWIZZLE_32B

    if (tid == 0) {
        for (int i = 0; i < NBOX; ++i) mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
    }

    // ---- tile start state: from parameter space (uniform)
    const int ka_rel = blockIdx.x * P.KT;                              // relative to k_begin
    const int ntile = min(P.KT, (int)(P.N - P.k_begin) - ka_rel);
    int j = P.tile[blockIdx.x].j;                                      // bank row (run order) of the next output
    int s = P.tile[blockIdx.x].s;                                      // window start, relative to box 0 of the tile
    const int xc0 = P.tile[blockIdx.x].xc0;                            // float coordinate of box 0 in tmx
    const int yc0 = ((int)P.k_begin + ka_rel) * 2;
    // boxes this tile is expected to touch (prefetch bound); demand may exceed it by a box or two
    const int jend = ((ntile + (int)(((long long)ntile * (P.M - P.L)) / P.L) + TPAD + RMAX) >> 3) + 1;
    __syncthreads();

    int k = 0;            // tile-relative index of the next output
    int j_issued = 0;     // boxes issued so far (tile-relative)
    int j_waited = 0;     // boxes already waited for
    int q_flushed = 0;    // output chunks already handed to TMA
    int pf_next = 0;      // next 8-box group to prefetch into L2 (thread 0 only)

    while (k < ntile) {
        const int rt = P.runtab[j];
        const int len = min(rt & 0xff, ntile - k);
        const int A = s & ~1;                        // aligned window start
        const int jA = A >> 3;                        // oldest live box
        const int jneed = (A + 2 * NPRUN - 1) >> 3;   // newest box the run's windows touch
        const int q_done = k >> 2;                    // chunks completed by earlier runs

        const bool flush = q_done > q_flushed;
        if (flush) fence_async_smem();                // make this thread's st.shared visible to the async proxy
        __syncthreads();                              // every warp finished the previous run (reads and writes)
        if (tid == 0) {
            // loads first: they are on the critical path of the NEXT run; the stores only have to leave eventually
            const int jtarget = max(jneed, min(jA + NBOX - 1, jend));
            for (int j = j_issued; j <= jtarget && j < jA + NBOX; ++j) {
                const uint32_t bar = bar_base + 8 * (j & (NBOX - 1));
                mbar_expect_tx(bar, BOX_BYTES);
                tma_load_2d(in_base + (uint32_t)((j & (NBOX - 1)) * BOX_BYTES), &tmx, xc0 + j * 16, ch0, bar);
            }
            if (P.pf_dist > 0) {
                // wide L2 prefetch ([64 ch][64 samples] = 512-byte rows) a few groups ahead of the ring loads:
                // DRAM sees long row bursts, the 64-byte-row box loads then hit L2
                for (; pf_next <= (jtarget >> 3) + P.pf_dist && pf_next * 8 <= jend; ++pf_next)
                    tma_prefetch_2d(&tmp, xc0 + pf_next * 128, ch0);
            }
            if (flush) {
                for (int q = q_flushed; q < q_done; ++q) {
                    tma_store_2d(&tmy, yc0 + q * 8, ch0, out_base + (uint32_t)((q & (kOutBufs - 1)) << 11));
                    tma_commit();
                }
                // Every store but the newest has finished reading its staging buffer.  A run completes at most 3
                // chunks and writes into at most 4, so with 8 buffers the compute warps never reach a buffer
                // whose store is still unconfirmed at the barrier above: no second barrier is needed.
                tma_wait_read<1>();
            }
        }
        {
            const int jtarget = max(jneed, min(jA + NBOX - 1, jend));
            j_issued = max(j_issued, min(jtarget, jA + NBOX - 1) + 1);
        }
        q_flushed = q_done;
        for (; j_waited <= jneed; ++j_waited)
            mbar_wait(bar_base + 8 * (j_waited & (NBOX - 1)), (uint32_t)((j_waited / NBOX) & 1));

        if (half == 0) {
            if (s & 1) run_body_c64<TPAD, RW, 0, 1, NBOX>(P, A, ibase_x, j, len, k, obase_x);
            else run_body_c64<TPAD, RW, 0, 0, NBOX>(P, A, ibase_x, j, len, k, obase_x);
        } else {
            if (s & 1) run_body_c64<TPAD, RW, RW, 1, NBOX>(P, A, ibase_x, j, len, k, obase_x);
            else run_body_c64<TPAD, RW, RW, 0, NBOX>(P, A, ibase_x, j, len, k, obase_x);
        }

        // ---- advance the (uniform) schedule by `len` outputs
        k += len;
        s += len + ((rt >> 8) & 1);                  // the run ended on a phase wrap: the input index skips one
        j = rt >> 16;
    }

    // ---- flush the remaining chunks (the last one may be partial: TMA clips at the tensor bound N)
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        const int q_end = (ntile + kOutChunk - 1) >> 2;
        for (int q = q_flushed; q < q_end; ++q) {
            tma_store_2d(&tmy, yc0 + q * 8, ch0, out_base + (uint32_t)((q & (kOutBufs - 1)) << 11));
            tma_commit();
        }
        tma_wait_read<0>();
    }
}

// =========================================================================================================
// k_tiled2_c64: same dataflow, re-cut for the memory system (tools/ubench3.cu: 128-byte rows copy 8 % faster
// than 64-byte rows, 256-byte L2 promotion reads 10 % faster) and for fewer instructions per output:
//  * CTA = 32 channels x 4 warps.  A ROUND is two consecutive runs; warp w computes half (w&1) of run (w>>1),
//    so one CTA barrier covers two runs.
//  * x boxes are [32 ch][16 samples] (128-byte rows, SWIZZLE_128B, 256-byte L2 promotion), 8-box ring = 128
//    samples: a round keeps <= 5 boxes live, so 3-4 boxes (12-16 KB per CTA) are always in flight.
//  * outputs are staged as [32 ch][16 outputs] (128-byte rows) in 4 buffers and leave by TMA store.
//  * every shared-memory address is  base + (per-lane constant XOR uniform word);  the uniform words come
//    from two small tables in parameter space (LDCU), so a window load is LOP3 + LDS.128 and nothing else.
// =========================================================================================================
constexpr int k2Rows = 32;              // channels per CTA
constexpr int k2BoxSamples = 16;        // samples per TMA box row (128 B for complex64)
constexpr int k2BoxBytes = k2Rows * k2BoxSamples * 8;   // 4096
constexpr int k2OutChunk = 16;          // outputs per staged TMA store (128-byte rows)
constexpr int k2OutBufs = 4;
constexpr int k2BankFloats = 5120;

struct alignas(16) Tiled2Params {
    long long k_begin, N;
    int L, M, KT, dbg;                                       // dbg: experiment switches (0 in production)
    struct Tile { int j, s, xc0, pad; } tile[kMaxTiles];     // as TiledParams::tile, boxes of 16 samples
    int runtab[kMaxPhases];                                  // as TiledParams::runtab
    // win[c][i] = W((c + i) mod 64):  W(u) = swizzle chunk bits | ring slot bits of sample pair u.  Four copies
    // shifted by c = 0..3 so that a window starting at any pair can be fetched with aligned 128-bit LDCU.
    unsigned win[4][80];
    // wout[i] = address word of staged output (i mod 64): chunk | odd/even | buffer
    unsigned wout[80];
    float bank[k2BankFloats];                                // run-ordered rows, as TiledParams::bank
};

template <int TPAD, int RW, int R0, int DELTA, int TW>
__device__ __forceinline__ void run_body2_c64(const Tiled2Params &P, const uint4 (&wq)[4], uint32_t in_base, uint32_t rowpart,
                                              int j, int len, int kpos, uint32_t out_base) {
    constexpr int NP = (TPAD + RW + 1) / 2;               // sample pairs this warp's outputs touch
    static_assert(NP <= 16, "four table quads cover the window");
    unsigned long long xw[2 * NP + 8];
    {
#pragma unroll
        for (int q = 0; q < NP; q += 4) {
            const unsigned w[4] = {wq[q / 4].x, wq[q / 4].y, wq[q / 4].z, wq[q / 4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (q + e < NP) {
                    const uint32_t a = in_base + (rowpart ^ w[e]);
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xw[2 * (q + e)]), "=l"(xw[2 * (q + e) + 1]) : "r"(a));
                }
            }
        }
    }
    const float *rows = P.bank + (j + R0) * TPAD;         // uniform base; everything below is base + constant
    len -= R0;
    const int o0 = (kpos + R0) & 63;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
#pragma unroll
        for (int i = 0; i < TPAD; i += 4) {
            float4 t;
            if constexpr (TW == 4) {
                t = *reinterpret_cast<const float4 *>(rows + r * TPAD + i);
            } else if constexpr (TW == 2) {
                const float2 t0 = *reinterpret_cast<const float2 *>(rows + r * TPAD + i);
                const float2 t1 = *reinterpret_cast<const float2 *>(rows + r * TPAD + i + 2);
                t = make_float4(t0.x, t0.y, t1.x, t1.y);
            } else {
                t = make_float4(rows[r * TPAD + i], rows[r * TPAD + i + 1], rows[r * TPAD + i + 2], rows[r * TPAD + i + 3]);
            }
            cfma(a0, t.x, xw[DELTA + r + i]);
            cfma(a1, t.y, xw[DELTA + r + i + 1]);
            cfma(a2, t.z, xw[DELTA + r + i + 2]);
            cfma(a3, t.w, xw[DELTA + r + i + 3]);
        }
        const unsigned long long y = cadd(cadd(a0, a1), cadd(a2, a3));
        if constexpr (TW == 1) {
            // one basic block per output: keeps ptxas from interleaving the rows of the bank (each row is one or two
            // constant-cache lines; six rows in flight per warp thrash the SM's small constant cache)
            const uint32_t a = out_base + (rowpart ^ P.wout[o0 + r]);
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(y) : "memory");
            if (r + 1 >= len) break;                      // uniform: the run is shorter than RW
        } else {
            if (r < len) {                                // uniform predicate
                const uint32_t a = out_base + (rowpart ^ P.wout[o0 + r]);
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(y) : "memory");
            }
        }
    }
}

template <int TPAD, int RMAX, int NBOX, int TW>
__global__ void __launch_bounds__(128, 4)
k_tiled2_c64(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
             const __grid_constant__ Tiled2Params P) {
    static_assert(NBOX == 8, "the address tables assume an 8-box ring");
    static_assert(RMAX % 4 == 0 && TPAD % 4 == 0, "runs split in two even halves, taps fetched four at a time");
    constexpr int RW = RMAX / 2;
    constexpr int NP = (TPAD + RW + 1) / 2;
    constexpr int WSPAN = RW + 2 * NP;                    // samples from a run's aligned start its two halves touch
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *in_ring = smem;
    unsigned char *out_ring = smem + NBOX * k2BoxBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(out_ring + k2OutBufs * k2Rows * k2OutChunk * 8);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction (keeps taps on LDCU)
    const int slot = warp >> 1, half = warp & 1;
    const bool issuer = tid == 96;                       // lane 0 of the warp that writes the newest outputs
    const int ch0 = blockIdx.y * k2Rows;
    const uint32_t in_base = smem_u32(in_ring), out_base = smem_u32(out_ring), bar_base = smem_u32(bars);
    // SWIZZLE_128B: 16-byte chunk index ^= row & 7
    const uint32_t rowpart = ((uint32_t)lane * 128u) ^ (((uint32_t)lane & 7u) << 4);

    if (tid == 0) {
        for (int i = 0; i < NBOX; ++i) mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
    }

    const int ka_rel = blockIdx.x * P.KT;
    const int ntile = min(P.KT, (int)(P.N - P.k_begin) - ka_rel);
    int j = P.tile[blockIdx.x].j;
    int s = P.tile[blockIdx.x].s;
    const int xc0 = P.tile[blockIdx.x].xc0;
    const int yc0 = ((int)P.k_begin + ka_rel) * 2;
    const int jend = ((ntile + (int)(((long long)ntile * (P.M - P.L)) / P.L) + TPAD + RMAX) >> 4) + 1;
    __syncthreads();

    int k = 0, j_issued = 0, j_waited = 0, q_flushed = 0;

    while (k < ntile) {
        // ---- the round's two runs (uniform)
        const int rtA = P.runtab[j];
        const int lenA = min(rtA & 0xff, ntile - k);
        const int sB = s + lenA + ((rtA >> 8) & 1);
        const int jB = rtA >> 16;
        const int kB = k + lenA;
        const int rtB = P.runtab[jB];
        const int lenB = min(rtB & 0xff, ntile - kB);
        const int jA = s >> 4;                              // oldest live box
        const int jneed = ((sB & ~1) + WSPAN - 1) >> 4;     // newest box the round's windows touch
        const int q_done = k >> 4;                          // chunks completed by earlier rounds

        // this warp's run of the round, and the address words of its window loads (fetched before the barrier so
        // that the loads can all issue the moment the data is there)
        const int myS = slot ? sB : s;
        const int myLen = slot ? lenB : lenA;
        uint4 wq[4];
        {
            const int p = (((myS & ~1) >> 1) + half * (RW / 2)) & 63;   // ring position (in pairs) of the window start
            const unsigned *wt = P.win[p & 3] + (p & ~3);
#pragma unroll
            for (int q = 0; q < 4; ++q) wq[q] = *reinterpret_cast<const uint4 *>(wt + 4 * q);
        }

        const bool flush = q_done > q_flushed;
        if (flush) fence_async_smem();
        if (issuer && !(P.dbg & 1)) tma_wait_read<0>();     // stores issued a round ago have left their buffers
        __syncthreads();
        const int jtarget = min(max(jneed, min(jA + NBOX - 1, jend)), jA + NBOX - 1);
        if (issuer) {
            for (int jj = j_issued; jj <= jtarget; ++jj) {
                const uint32_t bar = bar_base + 8 * (jj & (NBOX - 1));
                mbar_expect_tx(bar, k2BoxBytes);
                tma_load_2d(in_base + (uint32_t)((jj & (NBOX - 1)) * k2BoxBytes), &tmx, xc0 + jj * 32, ch0, bar);
            }
            if (flush && !(P.dbg & 4)) {
                for (int q = q_flushed; q < q_done; ++q) {
                    tma_store_2d(&tmy, yc0 + q * 32, ch0, out_base + (uint32_t)((q & (k2OutBufs - 1)) << 12));
                    tma_commit();
                }
                // this warp alone writes into the buffer of chunk q_done+2 == (q_done-2) mod 4 during the round
                if (!(P.dbg & 2)) tma_wait_read<1>();
            }
        }
        j_issued = max(j_issued, jtarget + 1);
        q_flushed = q_done;

        {
            // Every warp observes every box in order, busy or not: a parity wait is only meaningful while the
            // waiter is less than one ring revolution behind.
            const int need = ((myS & ~1) + half * RW + 2 * NP - 1) >> 4;
            for (; j_waited <= need; ++j_waited)
                mbar_wait(bar_base + 8 * (j_waited & (NBOX - 1)), (uint32_t)((j_waited >> 3) & 1));
        }
        if (myLen > half * RW && !(P.dbg & 8)) {
            const int myJ = slot ? jB : j, myK = slot ? kB : k;
            if (half == 0) {
                if (myS & 1) run_body2_c64<TPAD, RW, 0, 1, TW>(P, wq, in_base, rowpart, myJ, myLen, myK, out_base);
                else run_body2_c64<TPAD, RW, 0, 0, TW>(P, wq, in_base, rowpart, myJ, myLen, myK, out_base);
            } else {
                if (myS & 1) run_body2_c64<TPAD, RW, RW, 1, TW>(P, wq, in_base, rowpart, myJ, myLen, myK, out_base);
                else run_body2_c64<TPAD, RW, RW, 0, TW>(P, wq, in_base, rowpart, myJ, myLen, myK, out_base);
            }
        }

        k = kB + lenB;
        s = sB + lenB + ((rtB >> 8) & 1);
        j = rtB >> 16;
    }

    // ---- drain: every issued load must have landed before the CTA gives its shared memory back
    for (; j_waited < j_issued; ++j_waited)
        mbar_wait(bar_base + 8 * (j_waited & (NBOX - 1)), (uint32_t)((j_waited >> 3) & 1));
    fence_async_smem();
    __syncthreads();
    if (issuer) {
        const int q_end = (ntile + k2OutChunk - 1) >> 4;
        for (int q = q_flushed; q < q_end; ++q) {
            tma_store_2d(&tmy, yc0 + q * 32, ch0, out_base + (uint32_t)((q & (k2OutBufs - 1)) << 12));
            tma_commit();
        }
        tma_wait_read<0>();
    }
}

// =========================================================================================================
// k_tiled3_c64: shaped by the constant-cache measurement (profiles/README.md): taps fetched with LDCU are
// served by a small per-SM cache backed by the GPC constant cache (GCC), and the GCC sustains only ~4 requests
// per clock CHIP-WIDE.  Every CTA re-fetches each bank row once per L outputs, so GCC traffic is
// (channel groups) x (outputs) x (row bytes / 64) -- it falls only with the number of channels that share a
// fetch.  Hence:
//  * CTA = 128 channels x 4 warps; all four warps walk the SAME run at the same time (one fetch serves 128
//    channels), each computing the WHOLE run (<= 12 outputs) for its 32 channels from a 36-sample register
//    window: half the window loads per output of the split-run kernels, two code bodies instead of four.
//  * rows of the bank are consumed strictly one after the other (one basic block per output, next row's
//    taps prefetched into uniform registers during the current output) -- six rows in flight per warp
//    thrash the per-SM constant cache.
//  * x boxes are [128 ch][8 samples] (SWIZZLE_64B, 256-byte L2 promotion) in an 8-box ring; outputs are staged
//    as [128 ch][8 outputs] in 4 buffers and leave by TMA store.  2 CTAs per SM.
//  * every shared-memory address is base + (per-lane constant XOR uniform table word).
// =========================================================================================================
constexpr int k3Rows = 128;
constexpr int k3BoxSamples = 8;
constexpr int k3BoxBytes = k3Rows * k3BoxSamples * 8;   // 8192
constexpr int k3OutChunk = 8;
constexpr int k3OutBufs = 4;
constexpr int k3OutBytes = k3Rows * k3OutChunk * 8;     // 8192
constexpr int k3BankFloats = 5120;

struct alignas(16) Tiled3Params {
    long long k_begin, N;
    int L, M, KT, dbg;
    struct Tile { int j, s, xc0, pad; } tile[kMaxTiles];
    int runtab[kMaxPhases];
    unsigned win[4][64];      // win[c][i] = W((c + i) mod 32), W(u) = chunk bits | ring slot bits of sample pair u
    unsigned wout[48];        // wout[i] = address word of staged output (i mod 32)
    float4 bank[k3BankFloats / 4];   // declared float4 so that the rows are fetched with 128-bit uniform loads
};

// SPLIT = 1: every warp computes the whole run (4 warps per CTA); SPLIT = 2: the run is split over two warps
// (outputs 0..RMAX/2-1 and RMAX/2..RMAX-1; 8 warps per CTA) -- smaller windows, twice the warps per SM.
template <int NP>
__device__ __forceinline__ void load_window3(unsigned long long (&xw)[2 * NP], const unsigned *wt, uint32_t in_base, uint32_t rowpart) {
#pragma unroll
    for (int q = 0; q < NP; q += 4) {
        const uint4 w4 = *reinterpret_cast<const uint4 *>(wt + q);
        const unsigned w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (q + e < NP) {
                const uint32_t a = in_base + (rowpart ^ w[e]);
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xw[2 * (q + e)]), "=l"(xw[2 * (q + e) + 1]) : "r"(a) : "memory");
            }
        }
    }
}

template <int TPAD, int RW, int R0, int DELTA>
__device__ __forceinline__ void run_body3_c64(const Tiled3Params &P, const unsigned long long (&xw)[TPAD + RW + 1 - (TPAD + RW + 1) % 2],
                                              uint32_t rowpart, int j, int len, int kpos, uint32_t out_base) {
    const float4 *rows4 = P.bank + (j + R0) * (TPAD / 4);  // uniform base; everything below is base + constant
    const int o0 = (kpos + R0) & 31;
    len -= R0;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
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
        const uint32_t a = out_base + (rowpart ^ P.wout[o0 + r]);
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(y) : "memory");
        // uniform early exit = one basic block per output: the rows of the bank are then consumed strictly one
        // after the other (ptxas otherwise interleaves all rows of the run and thrashes the SM's constant cache)
        if (r + 1 >= len) break;
    }
}

template <int TPAD, int RMAX, int NBOX, int SPLIT>
__global__ void __launch_bounds__(128 * SPLIT, SPLIT == 1 ? 5 : 2)   // register cap: above ~100 ptxas moves taps to vector LDC
k_tiled3_c64(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
             const __grid_constant__ Tiled3Params P) {
    static_assert(NBOX == 8, "the address tables assume an 8-box ring");
    static_assert(TPAD % 4 == 0 && RMAX % (2 * SPLIT) == 0, "taps are fetched four at a time; halves start on even outputs");
    static_assert(7 + 2 * RMAX <= k3OutChunk * k3OutBufs, "staging buffers: see the store hazard note in the loop");
    constexpr int RW = RMAX / SPLIT;
    constexpr int NP = (TPAD + RW + 1) / 2;               // sample pairs one warp's outputs touch
    constexpr int NPRUN = (TPAD + RMAX + 1) / 2;          // sample pairs the whole run touches
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *in_ring = smem;
    unsigned char *out_ring = smem + NBOX * k3BoxBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(out_ring + k3OutBufs * k3OutBytes);

    const int tid = threadIdx.x;
    const int row = tid & (k3Rows - 1);
    const int half = SPLIT == 1 ? 0 : __shfl_sync(0xffffffffu, tid >> 7, 0);   // warp-uniform by construction
    const int ch0 = blockIdx.y * k3Rows;
    const uint32_t in_base = smem_u32(in_ring), out_base = smem_u32(out_ring), bar_base = smem_u32(bars);
    // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
    const uint32_t rowpart = ((uint32_t)row * 64u) ^ ((((uint32_t)row >> 1) & 3u) << 4);

    if (tid == 0) {
        for (int i = 0; i < NBOX; ++i) mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
    }

    const int ka_rel = blockIdx.x * P.KT;
    const int ntile = min(P.KT, (int)(P.N - P.k_begin) - ka_rel);
    int j = P.tile[blockIdx.x].j;
    int s = P.tile[blockIdx.x].s;
    const int xc0 = P.tile[blockIdx.x].xc0;
    const int yc0 = ((int)P.k_begin + ka_rel) * 2;
    const int jend = ((ntile + (int)(((long long)ntile * (P.M - P.L)) / P.L) + TPAD + RMAX) >> 3) + 1;
    if (tid == 0) {                                            // prologue: fill the ring
        for (int jj = 0; jj < NBOX; ++jj) {
            mbar_expect_tx(bar_base + 8 * jj, k3BoxBytes);
            tma_load_2d(in_base + (uint32_t)(jj * k3BoxBytes), &tmx, xc0 + jj * 16, ch0, bar_base + 8 * jj);
        }
    }
    __syncthreads();

    int k = 0, j_issued = NBOX, j_waited = 0, q_flushed = 0;

    while (k < ntile) {
        const int rt = P.runtab[j];
        const int len = min(rt & 0xff, ntile - k);
        const int A = s & ~1;
        const int jA = A >> 3;                                 // oldest live box
        const int jneed = (A + 2 * NPRUN - 1) >> 3;            // newest box the run's windows touch
        const int q_done = k >> 3;                             // chunks completed by earlier runs
        const int p = ((A >> 1) + half * (RW / 2)) & 31;       // ring position (in pairs) of this warp's window start
        const unsigned *wt = P.win[p & 3] + (p & ~3);

        // ---- this warp's window -> registers
        for (; j_waited <= jneed; ++j_waited)
            mbar_wait(bar_base + 8 * (j_waited & (NBOX - 1)), (uint32_t)((j_waited >> 3) & 1));
        unsigned long long xw[2 * NP];
        load_window3<NP>(xw, wt, in_base, rowpart);

        // ---- the one barrier of the run.  Behind it (a) every warp holds its window in registers, so all boxes
        // before the NEXT run's window are free and are refilled now, a whole run ahead of their first use; (b) every
        // warp has staged the previous run's outputs, so completed chunks can leave.
        const int s_next = s + len + ((rt >> 8) & 1);
        const int jA_next = (k + len < ntile) ? (s_next >> 3) : jA;
        const bool flush = q_done > q_flushed;
        if (flush) fence_async_smem();
        if (tid == 0) tma_wait_read<0>();                      // stores issued a run ago have left their buffers
        __syncthreads();
        const int jtarget = min(max(jneed, min(jA_next + NBOX - 1, jend)), jA_next + NBOX - 1);
        if (tid == 0) {
            for (int jj = j_issued; jj <= jtarget; ++jj) {
                const uint32_t bar = bar_base + 8 * (jj & (NBOX - 1));
                mbar_expect_tx(bar, k3BoxBytes);
                tma_load_2d(in_base + (uint32_t)((jj & (NBOX - 1)) * k3BoxBytes), &tmx, xc0 + jj * 16, ch0, bar);
            }
            // The run writes chunks q_done.. ; the stores issued here read the buffers of chunks q_flushed..q_done-1
            // (q_flushed = chunk of the previous run's first output).  Two runs span <= 7 + 2*RMAX = 31 outputs from
            // the start of chunk q_flushed, i.e. at most 4 chunks: with 4 buffers the two sets never share a
            // buffer, and everything older was confirmed by the wait above.
            if (flush) {
                for (int q = q_flushed; q < q_done; ++q) {
                    tma_store_2d(&tmy, yc0 + q * 16, ch0, out_base + (uint32_t)((q & (k3OutBufs - 1)) * k3OutBytes));
                    tma_commit();
                }
            }
        }
        j_issued = max(j_issued, jtarget + 1);
        q_flushed = q_done;

        if (len > half * RW && !(P.dbg & 8)) {
            if constexpr (SPLIT == 1) {
                if (s & 1) run_body3_c64<TPAD, RW, 0, 1>(P, xw, rowpart, j, len, k, out_base);
                else run_body3_c64<TPAD, RW, 0, 0>(P, xw, rowpart, j, len, k, out_base);
            } else {
                if (half == 0) {
                    if (s & 1) run_body3_c64<TPAD, RW, 0, 1>(P, xw, rowpart, j, len, k, out_base);
                    else run_body3_c64<TPAD, RW, 0, 0>(P, xw, rowpart, j, len, k, out_base);
                } else {
                    if (s & 1) run_body3_c64<TPAD, RW, RW, 1>(P, xw, rowpart, j, len, k, out_base);
                    else run_body3_c64<TPAD, RW, RW, 0>(P, xw, rowpart, j, len, k, out_base);
                }
            }
        }

        k += len;
        s = s_next;
        j = rt >> 16;
    }

    // ---- drain: every issued load must have landed before the CTA gives its shared memory back
    for (; j_waited < j_issued; ++j_waited)
        mbar_wait(bar_base + 8 * (j_waited & (NBOX - 1)), (uint32_t)((j_waited >> 3) & 1));
    fence_async_smem();
    if (tid == 0) tma_wait_read<0>();
    __syncthreads();
    if (tid == 0) {
        const int q_end = (ntile + k3OutChunk - 1) >> 3;
        for (int q = q_flushed; q < q_end; ++q) {
            tma_store_2d(&tmy, yc0 + q * 16, ch0, out_base + (uint32_t)((q & (k3OutBufs - 1)) * k3OutBytes));
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
    TiledParams *hp = nullptr;     // host template of the parameter block (bank + run lengths filled once)
    Tiled2Params *hp2 = nullptr;   // the same for k_tiled2_c64
    Tiled3Params *hp3 = nullptr;   // the same for k_tiled3_c64
    int variant = 3;               // 1 = k_tiled_c64 (64-channel CTAs), 2 = k_tiled2_c64 (32-channel CTAs, two runs per round)
    int kt_min = 1024;             // smallest time tile (outputs)
    int tw = 4;                    // tap fetch width (floats) of k_tiled2_c64
    PFN_encodeTiled encode = nullptr;
    int T = 0;
    int promo = 2;                 // L2 promotion of the x tensor map: 0 none, 1 128 B, 2 256 B
    std::vector<int> row_of_phase;
};

constexpr int kTPAD = 24, kRMAX = 12, kNBOX = 8;
constexpr int kTiledSmem = kNBOX * 4096 + kOutBufs * 2048 + 8 * kNBOX + 1024;
constexpr int kTiled3Smem = kNBOX * k3BoxBytes + k3OutBufs * k3OutBytes + 8 * kNBOX + 1024;
constexpr int kTiled2Smem = kNBOX * k2BoxBytes + k2OutBufs * k2Rows * k2OutChunk * 8 + 8 * kNBOX + 1024;

static inline void tiled_release(TiledPlan &p) {
    delete p.hp;
    delete p.hp2;
    delete p.hp3;
    p.hp3 = nullptr;
    p.hp = nullptr;
    p.hp2 = nullptr;
    p.ok = false;
}

// kind/tx/ty are the mrb.h enums (0 standard, 3 rational ; 2 = complex64)
static inline int32_t tiled_prepare(TiledPlan &p, int kind, int tx, int ty, int64_t L, int64_t M, int64_t Nphi,
                                    int64_t T, const std::vector<double> &bank, const std::vector<double> &,
                                    const cudaDeviceProp &) {
    p.ok = false;
    const bool kind_ok = kind == 0 /*standard*/ || kind == 3 /*rational*/;
    if (!kind_ok || tx != 2 || ty != 2) return 0;
    if (!(L <= M && M < 2 * L) || L > kMaxPhases || T > kTPAD || (L + kRMAX) * kTPAD > kBankFloats) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int32_t)(e ? e : cudaErrorUnknown);
    p.encode = (PFN_encodeTiled)fn;
    p.tpad = kTPAD; p.rmax = kRMAX; p.T = (int)T;
    p.hp = new TiledParams();
    memset(p.hp, 0, sizeof(TiledParams));
    p.hp->L = (int)L; p.hp->M = (int)M;
    if (const char *e = getenv("MRB_TILED_PF")) p.hp->pf_dist = atoi(e);
    if (const char *e = getenv("MRB_TILED_PROMO")) p.promo = atoi(e);
    if (const char *e = getenv("MRB_TILED_VARIANT")) p.variant = atoi(e);
    if (const char *e = getenv("MRB_TILED_TW")) p.tw = atoi(e);
    if (const char *e = getenv("MRB_TILED_KT")) p.kt_min = std::max(64, atoi(e) / 16 * 16);
    if ((L + kRMAX) * kTPAD > k2BankFloats) p.variant = 1;
    p.hp2 = new Tiled2Params();
    memset(p.hp2, 0, sizeof(Tiled2Params));
    p.hp2->L = (int)L; p.hp2->M = (int)M;
    if (const char *e = getenv("MRB_TILED_DBG")) p.hp2->dbg = atoi(e);
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 80; ++i) {
            const unsigned u = (unsigned)(c + i) & 63u;
            p.hp2->win[c][i] = ((u & 7u) << 4) | (((u >> 3) & 7u) << 12);
        }
    for (int i = 0; i < 80; ++i) {
        const unsigned kk = (unsigned)i & 63u;
        p.hp2->wout[i] = (((kk >> 1) & 7u) << 4) | ((kk & 1u) << 3) | (((kk >> 4) & 3u) << 12);
    }
    const int64_t mp = M - L;                                          // phase step per output
    p.row_of_phase.assign((size_t)L, 0);
    std::vector<int64_t> phase_of_row((size_t)L);
    for (int64_t j = 0; j < L; ++j) {
        phase_of_row[j] = (j * mp) % L;                                // a bijection: gcd(M-L, L) == gcd(M, L) == 1
        p.row_of_phase[phase_of_row[j]] = (int)j;
    }
    for (int64_t j = 0; j < L + kRMAX - 1; ++j) {
        const int64_t ph = phase_of_row[j % L];
        // row, left-padded with zeros: padded tap i multiplies the sample (TPAD-1-i) before the window's last one
        for (int64_t i = 0; i < T; ++i) p.hp->bank[j * kTPAD + (kTPAD - T) + i] = (float)bank[ph * T + i];
    }
    for (int64_t j = 0; j < L; ++j) {
        const int64_t ph = phase_of_row[j];
        // outputs at phases ph, ph+mp, ... read consecutive input windows until the phase wraps past L
        const int64_t to_wrap = mp == 0 ? kRMAX : (L - 1 - ph) / mp + 1;
        const int64_t len = std::min<int64_t>(to_wrap, kRMAX);
        const int64_t wrap = (mp != 0 && len == to_wrap) ? 1 : 0;
        p.hp->runtab[j] = (int)(len | (wrap << 8) | (((j + len) % L) << 16));
    }
    if (p.variant == 3) {
        p.hp3 = new Tiled3Params();
        memset(p.hp3, 0, sizeof(Tiled3Params));
        p.hp3->L = (int)L; p.hp3->M = (int)M; p.hp3->dbg = p.hp2->dbg;
        for (int c = 0; c < 4; ++c)
            for (int i = 0; i < 64; ++i) {
                const unsigned u = (unsigned)(c + i) & 31u;
                p.hp3->win[c][i] = ((u & 3u) << 4) | (((u >> 2) & 7u) << 13);
            }
        for (int i = 0; i < 48; ++i) {
            const unsigned kk = (unsigned)i & 31u;
            p.hp3->wout[i] = (((kk >> 1) & 3u) << 4) | ((kk & 1u) << 3) | (((kk >> 3) & 3u) << 13);
        }
        memcpy(p.hp3->runtab, p.hp->runtab, sizeof(p.hp->runtab));
        memcpy(reinterpret_cast<float *>(p.hp3->bank), p.hp->bank, sizeof(float) * (size_t)((L + kRMAX - 1) * kTPAD));
        e = cudaFuncSetAttribute(k_tiled3_c64<kTPAD, kRMAX, kNBOX, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiled3Smem);
        if (e != cudaSuccess) return (int32_t)e;
        e = cudaFuncSetAttribute(k_tiled3_c64<kTPAD, kRMAX, kNBOX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiled3Smem);
        if (e != cudaSuccess) return (int32_t)e;
    }
    if (p.variant == 2) {
        memcpy(p.hp2->runtab, p.hp->runtab, sizeof(p.hp->runtab));
        memcpy(p.hp2->bank, p.hp->bank, sizeof(float) * (size_t)((L + kRMAX - 1) * kTPAD));
        e = cudaFuncSetAttribute(k_tiled2_c64<kTPAD, kRMAX, kNBOX, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiled2Smem);
        if (e != cudaSuccess) return (int32_t)e;
        e = cudaFuncSetAttribute(k_tiled2_c64<kTPAD, kRMAX, kNBOX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiled2Smem);
        if (e != cudaSuccess) return (int32_t)e;
        e = cudaFuncSetAttribute(k_tiled2_c64<kTPAD, kRMAX, kNBOX, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiled2Smem);
        if (e != cudaSuccess) return (int32_t)e;

    }
    e = cudaFuncSetAttribute(k_tiled_c64<kTPAD, kRMAX, kNBOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiledSmem);
    if (e != cudaSuccess) return (int32_t)e;
    p.ok = true;
    return 0;
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

    if (p.variant == 3) {
        Tiled3Params &P = *p.hp3;
        P.k_begin = k_begin; P.N = G.nout;
        const int64_t span = G.nout - k_begin;
        P.KT = (int)std::max<int64_t>(p.kt_min, (ceil_div(span, kMaxTiles) + 15) / 16 * 16);
        const int64_t ntiles = ceil_div(span, P.KT);
        for (int64_t i = 0; i < ntiles; ++i) {
            const int64_t ka = k_begin + i * P.KT;
            const int64_t t0 = G.p0 + ka * G.M;
            const int64_t xs0 = G.d0m1 + t0 / G.L - (p.tpad - 1);     // x-sample index of the first window start
            const int64_t box0 = xs0 >> 3;
            P.tile[i].j = p.row_of_phase[(size_t)(t0 % G.L)];
            P.tile[i].s = (int)(xs0 - (box0 << 3));
            P.tile[i].xc0 = (int)(box0 << 3) * 2;
        }
        CUtensorMap tmx, tmy;
        cuuint64_t dims[2] = {(cuuint64_t)(2 * G.n_in), (cuuint64_t)G.nch};
        cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 8};
        cuuint32_t box[2] = {2 * k3BoxSamples, k3Rows};
        cuuint32_t es[2] = {1, 1};
        if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     p.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : p.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        cuuint64_t ydims[2] = {(cuuint64_t)(2 * G.nout), (cuuint64_t)G.nch};
        cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * 8};
        cuuint32_t ybox[2] = {2 * k3OutChunk, k3Rows};
        if (p.encode(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G.y, ydims, ystrides, ybox, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        dim3 grid((unsigned)ntiles, (unsigned)ceil_div(G.nch, k3Rows));
        if (p.tw == 1) k_tiled3_c64<kTPAD, kRMAX, kNBOX, 1><<<grid, 128, kTiled3Smem, st>>>(tmx, tmy, P);
        else k_tiled3_c64<kTPAD, kRMAX, kNBOX, 2><<<grid, 256, kTiled3Smem, st>>>(tmx, tmy, P);
        if (cudaPeekAtLastError() != cudaSuccess) return -2;
        *name = "tiled3_c64_t24_r12";
        ++*launches;
        return k_begin;
    }
    if (p.variant == 2) {
        Tiled2Params &P = *p.hp2;
        P.k_begin = k_begin; P.N = G.nout;
        const int64_t span = G.nout - k_begin;
        P.KT = (int)std::max<int64_t>(p.kt_min, (ceil_div(span, kMaxTiles) + 15) / 16 * 16);
        const int64_t ntiles = ceil_div(span, P.KT);
        for (int64_t i = 0; i < ntiles; ++i) {
            const int64_t ka = k_begin + i * P.KT;
            const int64_t t0 = G.p0 + ka * G.M;
            const int64_t xs0 = G.d0m1 + t0 / G.L - (p.tpad - 1);     // x-sample index of the first window start
            const int64_t box0 = xs0 >> 4;
            P.tile[i].j = p.row_of_phase[(size_t)(t0 % G.L)];
            P.tile[i].s = (int)(xs0 - (box0 << 4));
            P.tile[i].xc0 = (int)(box0 << 4) * 2;
        }
        CUtensorMap tmx, tmy;
        cuuint64_t dims[2] = {(cuuint64_t)(2 * G.n_in), (cuuint64_t)G.nch};
        cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 8};
        cuuint32_t box[2] = {2 * k2BoxSamples, k2Rows};
        cuuint32_t es[2] = {1, 1};
        if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     p.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : p.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        cuuint64_t ydims[2] = {(cuuint64_t)(2 * G.nout), (cuuint64_t)G.nch};
        cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * 8};
        cuuint32_t ybox[2] = {2 * k2OutChunk, k2Rows};
        if (p.encode(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G.y, ydims, ystrides, ybox, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        dim3 grid((unsigned)ntiles, (unsigned)ceil_div(G.nch, k2Rows));
        if (p.tw == 4) k_tiled2_c64<kTPAD, kRMAX, kNBOX, 4><<<grid, 128, kTiled2Smem, st>>>(tmx, tmy, P);
        else if (p.tw == 2) k_tiled2_c64<kTPAD, kRMAX, kNBOX, 2><<<grid, 128, kTiled2Smem, st>>>(tmx, tmy, P);
        else k_tiled2_c64<kTPAD, kRMAX, kNBOX, 1><<<grid, 128, kTiled2Smem, st>>>(tmx, tmy, P);
        if (cudaPeekAtLastError() != cudaSuccess) return -2;
        *name = "tiled2_c64_t24_r12";
        ++*launches;
        return k_begin;
    }
    TiledParams &P = *p.hp;
    P.k_begin = k_begin; P.N = G.nout;
    const int64_t span = G.nout - k_begin;
    P.KT = (int)std::max<int64_t>(p.kt_min, (ceil_div(span, kMaxTiles) + 15) / 16 * 16);
    const int64_t ntiles = ceil_div(span, P.KT);
    for (int64_t i = 0; i < ntiles; ++i) {
        const int64_t ka = k_begin + i * P.KT;
        const int64_t t0 = G.p0 + ka * G.M;
        const int64_t xs0 = G.d0m1 + t0 / G.L - (p.tpad - 1);         // x-sample index of the first window start
        const int64_t box0 = xs0 >> 3;
        P.tile[i].j = p.row_of_phase[(size_t)(t0 % G.L)];
        P.tile[i].s = (int)(xs0 - (box0 << 3));
        P.tile[i].xc0 = (int)(box0 << 3) * 2;
    }
    CUtensorMap tmx, tmy, tmp;
    {
        cuuint64_t dims[2] = {(cuuint64_t)(2 * G.n_in), (cuuint64_t)G.nch};
        cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 8};
        cuuint32_t box[2] = {2 * kBoxSamples, kTiledRows};
        cuuint32_t es[2] = {1, 1};
        if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     p.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : p.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        cuuint32_t pbox[2] = {16 * kBoxSamples, kTiledRows};
        if (p.encode(&tmp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, pbox, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        cuuint64_t ydims[2] = {(cuuint64_t)(2 * G.nout), (cuuint64_t)G.nch};
        cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * 8};
        cuuint32_t ybox[2] = {2 * kOutChunk, kTiledRows};
        if (p.encode(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G.y, ydims, ystrides, ybox, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
    }
    dim3 grid((unsigned)ntiles, (unsigned)ceil_div(G.nch, kTiledRows));
    k_tiled_c64<kTPAD, kRMAX, kNBOX><<<grid, kTiledThreads, kTiledSmem, st>>>(tmx, tmy, tmp, P);
    if (cudaPeekAtLastError() != cudaSuccess) return -2;
    *name = "tiled_c64_t24_r12";
    ++*launches;
    return k_begin;
}

}  // namespace mrb
