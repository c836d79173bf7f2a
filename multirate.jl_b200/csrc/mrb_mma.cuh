// mrb_mma.cuh -- tensor-core path (tcgen05, 3xTF32) for float32 samples x float32 taps.
//
// A polyphase FIR over many channels is a banded product.  For a GROUP of G consecutive outputs k = gG .. gG+G-1
//     Y[c, k] = sum_j X[c, a_g + j] * W_g[j, k - gG],       j = 0 .. K-1
// where a_g is the (8-sample aligned) start of the window the group's outputs share and column k of W_g holds the
// taps of output k (src/Filters.jl:284-298 bank rows; for the arbitrary-rate kinds the per-output row
// pfb[:,phi]+alpha*dpfb[:,phi] of :681-686 or the Farrow row of :789-791), shifted down by (window start of k) - a_g.
// With lane = channel this is D[128 channels x G] = A[128 x K] * B[K x G]: a real dense contraction (density T/K,
// 60-80 % at the BASELINE shapes), which the CUDA cores can only feed through shared-memory tap broadcasts
// (mrb_table.cuh: LSU bound at 24 % of FP32).  Here it runs on the 5th-generation tensor cores:
//
//  * kind::tf32 UMMA, M = 128 (channels), N = G, K = 8 per instruction, accumulators in TENSOR MEMORY;
//  * float32 accuracy through the 3xTF32 split  x = xh + xl, w = wh + wl,  y ~ xh*wh + xh*wl + xl*wh  (each part exactly
//    representable in tf32, round-to-nearest splits; the dropped xl*wl term is 2^-22 relative);
//  * A (the samples) lives in tensor memory as well: a RING OF COLUMNS, lane = channel, column = sample index mod RC.
//    Converter warps read every TMA box of x once from shared memory, split it and write xh / xl with tcgen05.st --
//    so a sample is converted once however many groups use it, shared memory holds only the boxes in flight, and the
//    MMAs read only B from shared memory (TS mode: SS mode would re-read the 128-row A tile for every instruction);
//  * B (the tap tiles, already split and stored as the K-major SWIZZLE_128B shared-memory image) is built once per
//    chunk for all channels by a pre-pass (k_mma_tiles) and streamed by bulk copies;
//  * warp-specialised, no CTA-wide barrier in the loop: x loader, tile loader, MMA issuer (one elected lane) with a look-ahead
//    warp that does its mbarrier waits one group ahead,
//    4 converter warps, 4 epilogue warps (tcgen05.ld -> swizzled staging -> TMA store), all coupled by mbarriers;
//    tcgen05.commit releases tile slots, ring columns and hands accumulators to the epilogue.
// Outputs whose window reaches into the history are computed by k_generic, as for every fast path.
#pragma once
#include <cstdio>

#include "mrb_tiled.cuh"

namespace mrb {

constexpr int kMmaRows = 128;           // channels per CTA = UMMA M
constexpr int kMmaBox = 32;             // samples per TMA box row (128 B) = one swizzle atom of K
constexpr int kMmaBoxBytes = kMmaRows * kMmaBox * 4;     // 16 KB
constexpr int kMmaNXB = 4;              // x boxes in flight in shared memory
constexpr int kMmaNAB = 7;              // boxes the tensor-memory rings hold (RC = 224 columns each)
constexpr int kMmaRC = kMmaNAB * kMmaBox;
constexpr int kMmaMaxKB = 6;            // K <= 192: the MMAs of a group read at most 7 boxes
constexpr int kMmaThreads = 384;        // warps 0-3 epilogue, 4-7 converters, 8 x loader, 9 tile loader, 10 MMA issuer, 11 its look-ahead waiter
constexpr int kMmaMaxGT = 512;          // groups per time tile (their window starts are staged in shared memory)

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// the mbarrier is arrived at once every MMA issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32
__device__ __forceinline__ void umma_ts_tf32(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 bytes apart, version 1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
// round to nearest (ties away) to tf32: the low 13 mantissa bits of the result are zero
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------------------
// pre-pass: the tap tiles of a schedule slice.  One warp per output row.
//   group g = k / G holds outputs gG .. gG+G-1; its window starts at gstart[g] = floor8(sn[gG] - H)
//   tile g = [hi | lo] x [KB atoms][G rows][32 floats], the K-major SWIZZLE_128B image the UMMA B descriptor reads:
//   element (row r, k index j) at  atom (j>>5) : r*128 + (((j&31)>>2 ^ (r&7)) << 4) + (j&3)*4  bytes
//   row r, index j = tap (j - d) of output gG+r,  d = (sn[k] - H) - gstart[g]   (zero outside the T taps)
// `mode`: 0 arbitrary (pfb + alpha*dpfb), 1 farrow (Horner), 2 integer schedule (branch sphi[k] of pfb, no blend)
// ---------------------------------------------------------------------------------------------------------
struct MmaSched {               // where the pre-pass takes an output's (window end, branch, blend) from
    int mode;                   // 0 arbitrary (pfb + alpha*dpfb), 1 farrow (Horner), 2 integer schedule (closed form)
    const int64_t *sn;          // modes 0/1: 0-based x index of the window's last sample, per output of the slice
    const int32_t *sphi;        // mode 0: 0-based branch
    const double *sa;           // mode 0: alpha; mode 1: Float64 phase
    long long L, M, p0, d0m1;   // mode 2: n_k = d0m1 + (p0 + k M) / L, branch (p0 + k M) % L   (k = slice-relative + k_base)
    long long k_base;
    // complex64 samples x real taps: the kernel runs on the FLOAT view of x and y (interleaved re, im).  Float output j = 2k + b
    // is output k's part b: its window ends at float 2 n_k + b and holds tap i of the branch at window position 2i (the odd
    // positions -- the other part's samples -- are zeros): a real FIR of 2T - 1 taps with the schedule below.  Same bytes per
    // float as the real kernel, twice the MMAs per output (the pipe has the room: 26 % active on the float32 resampler).
    int cplx;                   // 0 real; 1 complex64 on the float view (above); 2 complex64 SPLIT (below, k_mma_fir)
    long long span_c;           // cplx: widest spread of window starts inside a group of 16 outputs, in complex samples
};

__device__ __forceinline__ void mma_sched_at(const MmaSched &S, int64_t kf, int64_t &n, int64_t &phi) {
    const int64_t k = S.cplx == 1 ? kf >> 1 : kf;                      // output index (kf: float index of the interleaved view)
    if (S.mode == 2) {
        const long long t = S.p0 + (S.k_base + k) * S.M;
        const long long q = t / S.L;
        n = S.d0m1 + q;
        phi = t - q * S.L;
    } else {
        n = S.sn[k];
        phi = S.mode == 0 ? S.sphi[k] : 0;
    }
    if (S.cplx == 1) n = 2 * n + (kf & 1);
}

// Rows [0, nrows) of the tile table (nrows = tiles * G; for periodic integer schedules only one period of tiles is
// built) and the window starts of ALL ngroups groups.  `nout`: rows at or past it are zero (aperiodic tables only).
template <int G>
__global__ void __launch_bounds__(256)
k_mma_tiles(const float *__restrict__ pfb, const float *__restrict__ dpfb, const double *__restrict__ pnfb, int P1, int T, int KB,
            const MmaSched S, int64_t H, int64_t nout, int64_t nrows, int64_t ngroups, float *__restrict__ tiles,
            int32_t *__restrict__ gstart) {
    // split complex mode: a group is 16 (complex) outputs, the tile keeps its 32-row layout with rows 16..31 unused
    const int OG = S.cplx == 2 ? G / 2 : G;
    {   // window starts: one thread per group
        const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
        if (g < ngroups) {
            int64_t n, phi;
            mma_sched_at(S, g * OG, n, phi);
            const int64_t xg = n - H;
            gstart[g] = (int32_t)(xg >= 0 ? (xg & ~(int64_t)7) : -(((-xg) + 7) & ~(int64_t)7));
        }
    }
    const int64_t k = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= nrows) return;
    const int lane = threadIdx.x & 31;
    const int64_t g = k / G;
    const int r = (int)(k - g * G);
    int64_t ng0, phig;
    mma_sched_at(S, g * OG, ng0, phig);
    const int64_t xg = ng0 - H;
    const int64_t al = xg >= 0 ? (xg & ~(int64_t)7) : -(((-xg) + 7) & ~(int64_t)7);
    const int64_t ko = g * OG + r;                                   // output this row belongs to
    const bool live = r < OG && ko < nout;
    int64_t nk = 0, phik = 0;
    if (live) mma_sched_at(S, ko, nk, phik);
    const int d = live ? (int)(nk - H - al) : 0;
    const double ph = live && S.mode != 2 ? S.sa[S.cplx == 1 ? ko >> 1 : ko] : 0.0;   // farrow: phase; arbitrary: alpha
    const int Tw = S.cplx == 1 ? 2 * T - 1 : T;                      // window length in elements of the view
    const int64_t obase = phik * T;
    const int64_t tile_floats = (int64_t)2 * KB * G * 32;
    float *th = tiles + g * tile_floats, *tl = th + (int64_t)KB * G * 32;
    for (int j = lane; j < KB * 32; j += 32) {
        const int iw = j - d;                                        // window position
        const int i = S.cplx == 1 ? iw >> 1 : iw;                    // tap (float view: taps sit at the even positions)
        float v = 0.f;
        if (live && iw >= 0 && iw < Tw && !(S.cplx == 1 && (iw & 1))) {
            if (S.mode == 1) {
                // currentTaps[i] = polyval(pnfb[i], phase): Horner highest order first in Float64, separately rounded
                // multiply and add, rounded to the tap type (src/Filters.jl:789-791)
                const double *c = pnfb + (int64_t)i * P1;
                double a = c[P1 - 1];
                for (int p = P1 - 2; p >= 0; --p) a = __dadd_rn(__dmul_rn(a, ph), c[p]);
                v = (float)a;
            } else if (S.mode == 0) {
                v = (float)((double)pfb[obase + i] + ph * (double)dpfb[obase + i]);   // tapsforphase, :681-686
            } else {
                v = pfb[obase + i];                                                     // pfb[:, phi], :558-565
            }
        }
        const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
        const int off = (j >> 5) * (G * 32) + r * 32 + ((((j & 31) >> 2) ^ (r & 7)) << 2) + (j & 3);
        th[off] = hi;
        tl[off] = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------
struct alignas(16) MmaParams {
    long long g_begin, g_end;  // groups [g_begin, g_end) of the slice
    long long y0;              // output index of slice output 0 in y
    int GT;                    // groups per time tile
    int KB;                    // swizzle atoms (32 samples) per tile row: tile K extent = 32 KB
    int KS;                    // K-steps (8 samples) actually issued per group (<= 4 KB)
    int NWB;                   // tile slots in shared memory
    int tile_bytes;            // 2 * KB * G * 128
    int nissue;                // 2: a look-ahead warp does the issuer's waits (default); 1: the issuer waits itself (MRB_MMA_ISSUERS=1)
    int period;                // tile of group g = tile (g mod period): g_end - 0 for aperiodic tables
    int resident;              // period <= NWB: every tile is loaded once and stays in its slot
    int split;                 // 1: complex64 SPLIT mode.  Samples arrive as float boxes (re, im interleaved; 16 complex samples per box);
                               // the converters put the real parts and the imaginary parts into SEPARATE tensor-memory rings (16 columns
                               // per box each), so a group of 16 complex outputs is two N = 16 products with the SAME tap tile -- D_re =
                               // A_re W, D_im = A_im W -- instead of one N = 32 product on the float view whose tile is half zeros.  All
                               // window indices (gstart, KS, H) are in complex samples; the epilogue interleaves re / im again.
    int fwd;                   // 1: the epilogue releases tile slots and ring boxes once it sees the group's accumulators complete
                               // (one tcgen05.commit per group instead of three or four on the issuing thread); 0: MRB_MMA_FWD=0
    int nch;                   // channels (rows past it read zero history)
    const float *hist;         // [nch][H] history of the chunk: samples at x indices -H .. -1
    long long H;
    long long *prof;           // development aid (MRB_MMA_PROF=1): per CTA 16 counters of cycles spent waiting, see below
};

template <int G>
struct MmaCfg {
    static constexpr int OUT_BYTES = kMmaRows * G * 4;                // staging buffer of one group (G = 32: 128-byte rows)
    static constexpr int TM_D = 0;                                    // accumulators: 2 x G columns
    static constexpr int TM_AH = 2 * G;                               // hi ring
    static constexpr int TM_AL = 2 * G + kMmaRC;                      // lo ring
    static_assert(2 * G + 2 * kMmaRC <= 512, "tensor memory holds 512 columns");
    static_assert(G == 32, "the epilogue stages 128-byte rows");
};

// dynamic shared memory: x ring, tile ring, 2 staging buffers, window starts, 64 mbarrier slots, 1 KiB of alignment slack
static inline int mma_smem_fixed(int G) { return kMmaNXB * kMmaBoxBytes + 2 * kMmaRows * G * 4 + (kMmaMaxGT + 8) * 4 + 8 * 64 + 1024; }

// The MMAs of one group: KS K-steps of 8 samples, three MMAs each (xh*wh, xh*wl, xl*wh).  Straight-line
// code with a compile-time trip count: every K-step's operands are (group base) + (constant) and live in uniform
// registers of their own.  A rolled loop, and an unrolled one with a branch per K-step, both issued an MMA only every
// ~50 cycles -- each UTCHMMA holds its source uniform registers until the tensor pipe takes it, and the next step's
// address arithmetic was reusing them; the pipe takes one N = 32 MMA every 16 cycles (tools/mma_probe.cu, T6).
template <int G, int KS>
__device__ __forceinline__ void mma_issue_group(uint32_t d, uint32_t a_hi, uint32_t a_lo, int col0, uint64_t bh0, uint32_t lo_off,
                                                uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        int col = col0 + 8 * ks;                                       // 8 ks < RC: one conditional subtraction wraps the ring
        col -= col >= kMmaRC ? kMmaRC : 0;
        const uint64_t bh = bh0 + (uint64_t)(((uint32_t)((ks >> 2) * G * 128 + (ks & 3) * 32)) >> 4);
        const uint64_t bl = bh + lo_off;
        umma_ts_tf32(d, a_hi + (uint32_t)col, bh, idesc, ks > 0 ? 1u : 0u);
        umma_ts_tf32(d, a_hi + (uint32_t)col, bl, idesc, 1u);
        umma_ts_tf32(d, a_lo + (uint32_t)col, bh, idesc, 1u);
    }
}

// split complex mode: per K-step the same two tile halves serve the real-part ring and the imaginary-part ring (N = 16 each)
template <int G, int KS>
__device__ __forceinline__ void mma_issue_group_split(uint32_t d, uint32_t a_hi, uint32_t a_lo, int col0, uint64_t bh0, uint32_t lo_off,
                                                      uint32_t idesc) {
    constexpr int RCS = kMmaNAB * 16;                                  // columns of one part's ring
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        int col = col0 + 8 * ks;
        col -= col >= RCS ? RCS : 0;
        const uint64_t bh = bh0 + (uint64_t)(((uint32_t)((ks >> 2) * G * 128 + (ks & 3) * 32)) >> 4);
        const uint64_t bl = bh + lo_off;
#pragma unroll
        for (int part = 0; part < 2; ++part) {                         // 0: real parts, 1: imaginary parts
            const uint32_t dp = d + (uint32_t)(part * 16), c = (uint32_t)(part * RCS + col);
            umma_ts_tf32(dp, a_hi + c, bh, idesc, ks > 0 ? 1u : 0u);
            umma_ts_tf32(dp, a_hi + c, bl, idesc, 1u);
            umma_ts_tf32(dp, a_lo + c, bh, idesc, 1u);
        }
    }
}

// mbarrier wait that adds the cycles it spent to `acc` when profiling is on
__device__ __forceinline__ void mbar_wait_prof(uint32_t bar, uint32_t parity, bool prof, long long &acc) {
    if (!prof) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
}

template <int G, bool SPLIT>
__global__ void __launch_bounds__(kMmaThreads, 1)
k_mma_fir(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const float *__restrict__ tiles,
          const int32_t *__restrict__ gstart, const __grid_constant__ MmaParams P) {
    using C = MmaCfg<G>;
    constexpr int NXB = kMmaNXB, NAB = kMmaNAB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // dynamic shared memory is only guaranteed 16-byte aligned: the swizzled regions need 1 KiB
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *xring = smem;                                       // NXB boxes [128 ch][32 samples], SWIZZLE_128B
    unsigned char *wring = xring + NXB * kMmaBoxBytes;                 // NWB tiles
    unsigned char *oring = wring + P.NWB * P.tile_bytes;               // 2 staging buffers [128 ch][G outputs], SWIZZLE_128B
    int *gs_raw = reinterpret_cast<int *>(oring + 2 * C::OUT_BYTES);   // window starts of the tile's groups (16-byte granules)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(gs_raw + kMmaMaxGT + 8);
    uint32_t *tmem_base_p = reinterpret_cast<uint32_t *>(bars + 40);   // written by tcgen05.alloc

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ch0 = blockIdx.x * kMmaRows;
    const long long g0 = P.g_begin + (long long)blockIdx.y * P.GT;     // first group of the tile
    const int ng = (int)min((long long)P.GT, P.g_end - g0);
    const long long gal = g0 & ~3ll;                                   // bulk copies start on 16-byte boundaries
    const int *gs = gs_raw + (int)(g0 - gal);

    const uint32_t bar0 = smem_u32(bars);
    auto B_XFULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };                 // NXB
    auto B_XEMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(4 + i); };          // NXB
    auto B_AFULL = [&](int i) { return bar0 + 8u * (uint32_t)(8 + i); };           // NAB
    auto B_AEMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(16 + i); };         // NAB
    auto B_WFULL = [&](int i) { return bar0 + 8u * (uint32_t)(24 + i); };          // NWB <= 4
    auto B_WEMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(28 + i); };
    auto B_DFULL = [&](int i) { return bar0 + 8u * (uint32_t)(32 + i); };          // 2
    auto B_DEMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(34 + i); };
    const uint32_t B_GS = bar0 + 8u * 36u;
    auto B_GO = [&](int i) { return bar0 + 8u * (uint32_t)(42 + i); };             // 4 (slot 40 holds the tensor-memory base)

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int i = 0; i < NXB; ++i) { mbar_init(B_XFULL(i), 1); mbar_init(B_XEMPTY(i), 128); }
        for (int i = 0; i < NAB; ++i) { mbar_init(B_AFULL(i), 128); mbar_init(B_AEMPTY(i), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(B_WFULL(i), 1); mbar_init(B_WEMPTY(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(B_DFULL(i), 1); mbar_init(B_DEMPTY(i), 128); }
        mbar_init(B_GS, 1);
        for (int i = 0; i < 4; ++i) mbar_init(B_GO(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
        // the window starts of this tile's groups (whole 16-byte granules; the host allocates the slack)
        const uint32_t bytes = (uint32_t)(((int)(g0 - gal) + ng + 3) / 4 * 16);
        mbar_expect_tx(B_GS, bytes);
        bulk_load(smem_u32(gs_raw), gstart + gal, bytes, B_GS);
    }
    if (warp == 10) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_p)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_base_p;
    const bool prof = P.prof != nullptr;
    long long *pr = prof ? P.prof + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;                                  // cycles this role spent in its waits
    const long long t_start = prof ? clock64() : 0;

    mbar_wait(B_GS, 0);                                                // every role needs the window starts
    // sample index of box 0 of the tile.  It is negative for the tile that holds the chunk's first outputs, whose windows
    // reach into the history: boxes j < jneg lie wholly before x[0]; the converters fill them from the history buffer
    // (shiftin!'s carry, src/support.jl:61-80) instead of a TMA box, so no separate head kernel is needed.
    constexpr int BW = SPLIT ? kMmaBox / 2 : kMmaBox;                  // window samples (= ring columns) per box
    constexpr int BS = SPLIT ? 4 : 5;                                  // log2(BW)
    const int xbase = gs[0] & ~(BW - 1);
    const int jneg = xbase < 0 ? (-xbase) >> BS : 0;
    const int jlast = (gs[ng - 1] - xbase + P.KS * 8 - 1) >> BS;       // newest box the tile reads

    if (warp == 8) {
        // ---------------- x loader: TMA boxes into the shared-memory ring
        if (elect_one()) {
            for (int jx = 0; jx + jneg <= jlast; ++jx) {                // jx counts the boxes that come from x
                const int s = jx % NXB;
                if (jx >= NXB) mbar_wait_prof(B_XEMPTY(s), (uint32_t)((jx / NXB - 1) & 1), prof, c0);
                mbar_expect_tx(B_XFULL(s), kMmaBoxBytes);
                tma_load_2d(smem_u32(xring) + (uint32_t)(s * kMmaBoxBytes), &tmx, (xbase + (jx + jneg) * BW) * (SPLIT ? 2 : 1), ch0, B_XFULL(s));
            }
            if (prof) pr[0] = c0;
        }
    } else if (warp == 9) {
        // ---------------- tile loader: one bulk copy per group; a periodic schedule with few distinct tiles (standard,
        // interpolator: one) keeps them all resident instead
        if (elect_one()) {
            if (P.resident) {
                for (int t = 0; t < P.period; ++t) {
                    mbar_expect_tx(B_WFULL(t), (uint32_t)P.tile_bytes);
                    bulk_load(smem_u32(wring) + (uint32_t)(t * P.tile_bytes),
                              reinterpret_cast<const unsigned char *>(tiles) + (size_t)t * (size_t)P.tile_bytes, (uint32_t)P.tile_bytes, B_WFULL(t));
                }
            } else {
                int ti = (int)(g0 % P.period);
                for (int w = 0; w < ng; ++w) {
                    const int s = w % P.NWB;
                    if (w >= P.NWB) mbar_wait_prof(B_WEMPTY(s), (uint32_t)((w / P.NWB - 1) & 1), prof, c0);
                    mbar_expect_tx(B_WFULL(s), (uint32_t)P.tile_bytes);
                    bulk_load(smem_u32(wring) + (uint32_t)(s * P.tile_bytes),
                              reinterpret_cast<const unsigned char *>(tiles) + (size_t)ti * (size_t)P.tile_bytes, (uint32_t)P.tile_bytes, B_WFULL(s));
                    if (++ti == P.period) ti = 0;
                }
            }
            if (prof) pr[1] = c0;
        }
    } else if (warp == 11) {
        // ---------------- look-ahead waiter.  Everything group w must wait for -- its window converted (a_full), its tile
        // landed (w_full), its accumulator buffer drained (d_empty) -- is waited for HERE, one group ahead of the issuer,
        // and folded into one mbarrier GO[w & 3].  Every try_wait costs ~90 cycles even when it succeeds at once; with the
        // issuer doing the three or four of them itself, plus fence and commits, the tensor pipe sat idle ~700 cycles per
        // group (41 % tensor-pipe activity, profiles/r2_ncu_mma_c3b).  (Two issuer warps alternating groups hid the same
        // latency and measured +23 %, but faulted about once in 1000 launches: MMAs are issued by ONE thread only.)
        if (P.nissue == 2) {
            int boxes_ready = 0, ar_slot = 0;
            uint32_t ar_par = 0, w_par = 0;
            const int ws_wrap = P.resident ? P.period : P.NWB;
            int ws = P.resident ? (int)(g0 % P.period) : 0;
            for (int w = 0; w < ng; ++w) {
                const int need = (gs[w] - xbase + P.KS * 8 - 1) >> BS;
                for (; boxes_ready <= need; ++boxes_ready) {
                    mbar_wait(B_AFULL(ar_slot), ar_par);
                    if (++ar_slot == NAB) { ar_slot = 0; ar_par ^= 1u; }
                }
                mbar_wait(B_WFULL(ws), w_par);
                // (d_empty of group w also proves that the issuer has consumed GO of group w - 2, hence of w - 4: the
                // four GO barriers are never lapped)
                if (w >= 2) mbar_wait(B_DEMPTY(w & 1), (uint32_t)(((w >> 1) - 1) & 1));
                __syncwarp();
                if (lane == 0) mbar_arrive(B_GO(w & 3));
                if (++ws == ws_wrap) { ws = 0; if (!P.resident) w_par ^= 1u; }
            }
        }
    } else if (warp == 10) {
        // ---------------- MMA issuer (one elected lane)
        const bool ahead = P.nissue == 2;                              // the waits are done by the look-ahead warp
        const uint32_t idesc = umma_idesc_tf32(kMmaRows, SPLIT ? G / 2 : G);
        const uint32_t lo_off = (uint32_t)(P.KB * G * 128) >> 4;       // hi -> lo half of a tile, in descriptor units
        int boxes_ready = 0, dead = 0;
        int ar_slot = 0;                                               // a_full slot of box `boxes_ready` and its phase parity
        uint32_t ar_par = 0;
        const int ws_wrap = P.resident ? P.period : P.NWB;
        int ws = P.resident ? (int)(g0 % P.period) : 0;                 // tile slot of group w and its parity, kept without divisions
        uint32_t w_par = 0;
        long long c4 = 0;
        for (int w = 0; w < ng; ++w) {
            const int a0 = gs[w] - xbase;                              // multiple of 8
            const int need = (a0 + P.KS * 8 - 1) >> BS;
            // ring boxes no later group reads: everything before the next group's first box
            const int next_first = w + 1 < ng ? (gs[w + 1] - xbase) >> BS : jlast + 1;
            if (ahead) {
                mbar_wait_prof(B_GO(w & 3), (uint32_t)((w >> 2) & 1), prof, c0);
            } else {
                for (; boxes_ready <= need; ++boxes_ready) {
                    mbar_wait_prof(B_AFULL(ar_slot), ar_par, prof, c0);
                    if (++ar_slot == NAB) { ar_slot = 0; ar_par ^= 1u; }
                }
                mbar_wait_prof(B_WFULL(ws), w_par, prof, c1);
                if (w >= 2) mbar_wait_prof(B_DEMPTY(w & 1), (uint32_t)(((w >> 1) - 1) & 1), prof, c2);
            }
            const long long tf0 = prof ? clock64() : 0;
            tc_fence_after();
            if (prof) c4 += clock64() - tf0;
            const uint32_t d = tb + (uint32_t)(C::TM_D + (w & 1) * G);
            const uint64_t bh0 = umma_desc_sw128(smem_u32(wring) + (uint32_t)(ws * P.tile_bytes));
            const int col0 = SPLIT ? a0 % (NAB * 16) : a0 % kMmaRC;   // (constant divisors)
            const long long ti0 = prof ? clock64() : 0;
            if (elect_one()) {
                const uint32_t ah = tb + (uint32_t)C::TM_AH, al = tb + (uint32_t)C::TM_AL;
                if constexpr (SPLIT) {
                    switch (P.KS) {                                      // (the host admits K <= 96 complex columns: 12 K-steps)
                    case 1: mma_issue_group_split<G, 1>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 2: mma_issue_group_split<G, 2>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 3: mma_issue_group_split<G, 3>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 4: mma_issue_group_split<G, 4>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 5: mma_issue_group_split<G, 5>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 6: mma_issue_group_split<G, 6>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 7: mma_issue_group_split<G, 7>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 8: mma_issue_group_split<G, 8>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 9: mma_issue_group_split<G, 9>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 10: mma_issue_group_split<G, 10>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    case 11: mma_issue_group_split<G, 11>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    default: mma_issue_group_split<G, 12>(d, ah, al, col0, bh0, lo_off, idesc); break;
                    }
                } else
                switch (P.KS) {                                          // compile-time trip counts: straight-line issue code
                case 1: mma_issue_group<G, 1>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 2: mma_issue_group<G, 2>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 3: mma_issue_group<G, 3>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 4: mma_issue_group<G, 4>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 5: mma_issue_group<G, 5>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 6: mma_issue_group<G, 6>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 7: mma_issue_group<G, 7>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 8: mma_issue_group<G, 8>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 9: mma_issue_group<G, 9>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 10: mma_issue_group<G, 10>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 11: mma_issue_group<G, 11>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 12: mma_issue_group<G, 12>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 13: mma_issue_group<G, 13>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 14: mma_issue_group<G, 14>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 15: mma_issue_group<G, 15>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 16: mma_issue_group<G, 16>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 17: mma_issue_group<G, 17>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 18: mma_issue_group<G, 18>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 19: mma_issue_group<G, 19>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 20: mma_issue_group<G, 20>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 21: mma_issue_group<G, 21>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 22: mma_issue_group<G, 22>(d, ah, al, col0, bh0, lo_off, idesc); break;
                case 23: mma_issue_group<G, 23>(d, ah, al, col0, bh0, lo_off, idesc); break;
                default: mma_issue_group<G, 24>(d, ah, al, col0, bh0, lo_off, idesc); break;
                }
                // The issuing thread is the kernel's bottleneck (per group: 20 cycles per MMA + ~320 of fixed cost, most of it
                // commits): ONE commit hands the accumulators over, the epilogue -- which then knows that every MMA of the group
                // is done -- releases the tile slot and the dead ring boxes with plain arrivals (P.fwd).
                if (!P.fwd && !P.resident) tc_commit(B_WEMPTY(ws));    // the tile slot may be refilled (resident: never refilled)
                tc_commit(B_DFULL(w & 1));                             // the accumulators are complete
                if (!P.fwd)
                    for (int b = dead; b < next_first; ++b) tc_commit(B_AEMPTY(b % NAB));   // ring boxes nobody reads any more
            }
            __syncwarp();
            if (prof) c3 += clock64() - ti0;
            if (next_first > dead) dead = next_first;
            if (++ws == ws_wrap) { ws = 0; if (!P.resident) w_par ^= 1u; }   // resident tiles: phase 0 stays complete
        }
        if (prof && lane == 0) { pr[15] = c4; pr[2] = c0; pr[3] = c1; pr[4] = c2; pr[11] = c3; pr[14] = clock64() - t_start; }
    } else if (warp >= 4) {
        // ---------------- converters: thread = channel = tensor-memory lane; box j -> columns (j mod NAB) * 32 of both rings
        const int q = warp - 4;                                        // lane quadrant (warp index mod 4)
        const int row = q * 32 + lane;
        const uint32_t lanebase = tb + ((uint32_t)(q * 32) << 16);
        const uint32_t rowpart = ((uint32_t)row * 128u) ^ (((uint32_t)row & 7u) << 4);     // SWIZZLE_128B
        const bool rowlive = ch0 + row < P.nch;
        const float *hrow = P.hist + (long long)(ch0 + row) * P.H * (SPLIT ? 2 : 1);   // (split: P.H counts complex samples)
        for (int j = 0; j <= jlast; ++j) {
            const int jx = j - jneg;                                   // >= 0: box jx of the x ring; < 0: a history box
            const int s = jx >= 0 ? jx % NXB : 0, as = j % NAB;
            if (jx >= 0) mbar_wait_prof(B_XFULL(s), (uint32_t)((jx / NXB) & 1), prof, c0);
            if (j >= NAB) {
                mbar_wait_prof(B_AEMPTY(as), (uint32_t)((j / NAB - 1) & 1), prof, c1);
                tc_fence_after();
            }
            const uint32_t src = smem_u32(xring) + (uint32_t)(s * kMmaBoxBytes);
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                float x[16];
                if (jx >= 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t ad = src + (rowpart ^ ((uint32_t)(hlf * 4 + c) << 4));
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(x[4 * c]), "=f"(x[4 * c + 1]), "=f"(x[4 * c + 2]), "=f"(x[4 * c + 3]) : "r"(ad) : "memory");
                    }
                } else {
                    // samples n = xbase + 32 j + 16 hlf + e < 0: ext index H + n of [history | x]; zero before the history
                    if constexpr (SPLIT) {                                     // float e of the half box = part e & 1 of complex sample n0 + e / 2
                        const int n0 = xbase + j * BW + hlf * 8;
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const long long hi = P.H + n0 + (e >> 1);
                            x[e] = (rowlive && hi >= 0) ? __ldg(hrow + 2 * hi + (e & 1)) : 0.f;
                        }
                    } else {
                        const int n0 = xbase + j * kMmaBox + hlf * 16;
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const long long hi = P.H + n0 + e;
                            x[e] = (rowlive && hi >= 0) ? __ldg(hrow + hi) : 0.f;
                        }
                    }
                }
                uint32_t vh[16], vl[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float h = tf32_rna(x[e]);
                    vh[e] = __float_as_uint(h);
                    vl[e] = __float_as_uint(tf32_rna(x[e] - h));
                }
                if constexpr (SPLIT) {
                    // real parts -> columns [0, 112) of each ring, imaginary parts -> [112, 224): 8 complex samples per half box
                    uint32_t rh[8], ih[8], rl[8], il[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { rh[e] = vh[2 * e]; ih[e] = vh[2 * e + 1]; rl[e] = vl[2 * e]; il[e] = vl[2 * e + 1]; }
                    const uint32_t cc = (uint32_t)(as * 16 + hlf * 8);
                    tmem_st8(lanebase + (uint32_t)C::TM_AH + cc, rh);
                    tmem_st8(lanebase + (uint32_t)C::TM_AH + (uint32_t)(NAB * 16) + cc, ih);
                    tmem_st8(lanebase + (uint32_t)C::TM_AL + cc, rl);
                    tmem_st8(lanebase + (uint32_t)C::TM_AL + (uint32_t)(NAB * 16) + cc, il);
                } else {
                    tmem_st16(lanebase + (uint32_t)(C::TM_AH + as * kMmaBox + hlf * 16), vh);
                    tmem_st16(lanebase + (uint32_t)(C::TM_AL + as * kMmaBox + hlf * 16), vl);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(B_AFULL(as));
            if (jx >= 0) mbar_arrive(B_XEMPTY(s));
        }
        if (prof && tid == 128) { pr[5] = c0; pr[6] = c1; pr[13] = clock64() - t_start; }
    } else {
        // ---------------- epilogue: warp q reads tensor-memory lanes 32q .. 32q+31 (its channels), stages, stores
        const int row = warp * 32 + lane;
        const uint32_t lanebase = tb + ((uint32_t)(warp * 32) << 16);
        const uint32_t rowpart = ((uint32_t)row * 128u) ^ (((uint32_t)row & 7u) << 4);
        int dead_e = 0, ws_e = 0;                                      // (warp 1 lane 0) boxes and tile slot released so far
        for (int w = 0; w < ng; ++w) {
            mbar_wait_prof(B_DFULL(w & 1), (uint32_t)((w / 2) & 1), prof, c0);
            tc_fence_after();
            if (P.fwd && tid == 32) {
                // every MMA of group w has completed: its tile slot and the ring boxes before the next group's window are free
                if (!P.resident) { mbar_arrive(B_WEMPTY(ws_e)); if (++ws_e == P.NWB) ws_e = 0; }
                const int next_first = w + 1 < ng ? (gs[w + 1] - xbase) >> BS : jlast + 1;
                for (int b = dead_e; b < next_first; ++b) mbar_arrive(B_AEMPTY(b % NAB));
                if (next_first > dead_e) dead_e = next_first;
            }
            uint32_t v[G];
            tmem_ld16(lanebase + (uint32_t)(C::TM_D + (w & 1) * G), v);
            tmem_ld16(lanebase + (uint32_t)(C::TM_D + (w & 1) * G + 16), v + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(B_DEMPTY(w & 1));
            if constexpr (SPLIT) {                                             // columns 0..15 real parts, 16..31 imaginary parts -> interleaved
                uint32_t t[G];
#pragma unroll
                for (int i = 0; i < G / 2; ++i) { t[2 * i] = v[i]; t[2 * i + 1] = v[G / 2 + i]; }
#pragma unroll
                for (int i = 0; i < G; ++i) v[i] = t[i];
            }
            // the staging buffer of group w-2 must have been read by its TMA store
            if (tid == 0) {
                const long long t0 = prof ? clock64() : 0;
                tma_wait_read<1>();
                if (prof) c1 += clock64() - t0;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const uint32_t ob = smem_u32(oring) + (uint32_t)((w & 1) * C::OUT_BYTES);
#pragma unroll
            for (int c = 0; c < G / 4; ++c) {
                const uint32_t ad = ob + (rowpart ^ ((uint32_t)c << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(v[4 * c]), "r"(v[4 * c + 1]), "r"(v[4 * c + 2]),
                             "r"(v[4 * c + 3]) : "memory");
            }
            fence_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) {
                tma_store_2d(&tmy, (int)(P.y0 + (g0 + w) * G), ch0, ob);
                tma_commit();
            }
        }
        if (tid == 0) tma_wait_read<0>();
        if (prof && tid == 0) { pr[7] = c0; pr[8] = c1; pr[12] = clock64() - t_start; pr[9] = ng; pr[10] = jlast + 1; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 10) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
constexpr int kMmaG = 32;

struct MmaRows {                       // per pipeline stream (TableCtx): the tap tiles and window starts of one slice
    float *d_tiles = nullptr;
    int32_t *d_gstart = nullptr;
    int64_t cap_tiles = 0, cap_groups = 0;
    int64_t tile_bytes = 0;
    uint64_t tag = 0;                  // (call serial, slice) the tiles were built for: 0 = none
};

static inline void mmarows_release(MmaRows &r) {
    cudaFree(r.d_tiles); cudaFree(r.d_gstart);
    r = MmaRows{};
}

struct MmaPlan {
    bool ok = false;
    int cplx = 0;                      // complex64 samples: every launch works on the float view (see MmaSched::cplx)
    int T = 0;
    PFN_encodeTiled encode = nullptr;
    int num_sms = 148;
    int max_smem = 0;
};

static inline void mma_release(MmaPlan &p) { p.ok = false; }

// kind/tx/ty/th are the mrb.h enums (0 standard, 1 interpolator, 2 decimator, 3 rational, 4 arbitrary, 5 farrow; 0 = float32)
static inline int32_t mma_prepare(MmaPlan &p, int kind, int tx, int ty, int th, int64_t T, const cudaDeviceProp &prop) {
    p.ok = false;
    static const bool off = getenv("MRB_NO_MMA") != nullptr;
    static const bool no_c64 = getenv("MRB_MMA_C64") && atoi(getenv("MRB_MMA_C64")) == 0;
    // float32 samples, or complex64 samples through their float view (tx / ty: 0 = float32, 2 = complex64); float32 taps
    const bool real = tx == 0 && ty == 0, cplx = tx == 2 && ty == 2 && !no_c64;
    if (off || kind < 0 || kind > 5 || !(real || cplx) || th != 0) return 0;
    p.cplx = cplx ? 1 : 0;
    if ((cplx ? 2 * T - 1 : T) + 7 > kMmaMaxKB * 32) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int32_t)(e ? e : cudaErrorUnknown);
    p.encode = (PFN_encodeTiled)fn;
    p.num_sms = prop.multiProcessorCount;
    p.T = (int)T;
    p.max_smem = (int)prop.sharedMemPerBlockOptin;
    e = cudaFuncSetAttribute(k_mma_fir<kMmaG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.max_smem);
    if (e != cudaSuccess) return (int32_t)e;
    e = cudaFuncSetAttribute(k_mma_fir<kMmaG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.max_smem);
    if (e != cudaSuccess) return (int32_t)e;
    p.ok = true;
    return 0;
}

static inline cudaError_t mma_reserve(MmaRows &r, int64_t ntiles, int64_t groups, int64_t tile_bytes) {
    if (r.cap_tiles >= ntiles && r.cap_groups >= groups && r.tile_bytes == tile_bytes) return cudaSuccess;
    mmarows_release(r);
    cudaError_t e = cudaMalloc(&r.d_tiles, (size_t)ntiles * (size_t)tile_bytes);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&r.d_gstart, (size_t)(groups + 8) * sizeof(int32_t));
    if (e != cudaSuccess) return e;
    r.cap_tiles = ntiles; r.cap_groups = groups; r.tile_bytes = tile_bytes;
    return cudaSuccess;
}

// One schedule slice: outputs [0, cnt) of the slice (y index y0 + k).  `S` says where the schedule comes from (uploaded
// arrays for the arbitrary-rate kinds, the closed form for the integer kinds); max_group_span = widest spread of window
// starts inside a group of kMmaG outputs.  Builds the tap tiles, then launches the main kernel for the whole slice
// (chunk head included).  Returns 0 (the first output covered), -1 when not covered, -2 on a CUDA error.
static inline int64_t mma_try_launch(MmaPlan &p, MmaRows &rw, const GenParams &G, const MmaSched &S_in, int P1,
                                     const void *d_pfb, const void *d_dpfb, const double *d_pnfb, int64_t y0, int64_t cnt,
                                     int64_t max_group_span, cudaStream_t st, const char **name, int64_t *launches, uint64_t tag = 0) {
    static const bool trace = getenv("MRB_TRACE") != nullptr;
#define MRB_MMA_SKIP(why) do { if (trace) fprintf(stderr, "[mrb] tensor-core kernel not used: %s\n", why); return -1; } while (0)
    if (!p.ok) MRB_MMA_SKIP("configuration not covered");
    constexpr int GG = kMmaG;
    if (((uintptr_t)G.x & 15) || ((uintptr_t)G.y & 15) || (G.ldx % 4) || (G.ldy % 4)) MRB_MMA_SKIP("alignment");
    if (G.n_in >= (1ll << 31) - 4096 || y0 + cnt >= (1ll << 31) - 4096) MRB_MMA_SKIP("size");
    if (y0 % GG) MRB_MMA_SKIP("slice start");
    // a CTA is 128 channels wide (UMMA M): with a handful of channels the other kernels win (README benchmark, ONE channel:
    // k_stream 0.024 ms, this kernel 0.19 ms)
    if (G.nch < 48) MRB_MMA_SKIP("too few channels for a 128-row tile");
    // complex64: the SPLIT form (real and imaginary parts in separate tensor-memory rings of 7 x 16 columns, half the MMA work)
    // when a group's window fits 96 complex columns, else the float view (cplx = 1)
    static const bool no_split = getenv("MRB_MMA_SPLIT") && atoi(getenv("MRB_MMA_SPLIT")) == 0;
    MmaSched S = S_in;
    const bool split = S.cplx == 1 && !no_split && (int64_t)p.T + 7 + S.span_c <= 96;
    if (split) { S.cplx = 2; max_group_span = S.span_c; }
    const int64_t Tw = S.cplx == 1 ? 2 * (int64_t)p.T - 1 : p.T;      // window length in elements of the view
    const int64_t kneed = Tw + 7 + max_group_span;                    // samples a group's window spans at worst
    const int KB = (int)ceil_div(kneed, 32);
    if (KB > kMmaMaxKB) MRB_MMA_SKIP("window group wider than the tensor-memory ring");
    const int KS = (int)ceil_div(kneed, 8);
    // the kernel covers the whole slice: windows that reach into the history are filled from the history buffer
    const int64_t k_begin = 0;
    const int64_t groups = ceil_div(cnt, GG), g_begin = 0;
    if (groups < 8) MRB_MMA_SKIP("slice too short");
    if (G.H > (1ll << 20)) MRB_MMA_SKIP("history too long");
    const int tile_bytes = 2 * KB * GG * 128;
    const int fixed = mma_smem_fixed(GG);
    const int nwb = std::min(4, (p.max_smem - fixed) / tile_bytes);
    if (nwb < 2) MRB_MMA_SKIP("shared memory");
    // Integer schedules repeat: group g + P reads the same taps at the same alignment as group g once P groups advance the
    // phase by a multiple of L and the input by a multiple of 8 samples.  Only one period of tiles is built (standard and
    // interpolators: ONE tile, kept resident in shared memory; 147//160: 147 tiles, L2 resident).
    int64_t period = groups;
    if (S.mode == 2) {
        const int64_t og = S.cplx ? GG / 2 : GG, adv = S.cplx == 1 ? 2 : 1;  // outputs per group; view elements per input sample
        for (int64_t q = 1; q <= std::min<int64_t>(groups, 16 * S.L); ++q)
            if ((q * og * S.M) % S.L == 0 && (((q * og * S.M) / S.L) * adv) % 8 == 0) { period = q; break; }
    }
    const int64_t ntiles = std::min(period, groups);
    const bool resident = S.mode == 2 && period <= nwb;
    {
        const float *before = rw.d_tiles;
        if (mma_reserve(rw, ntiles, groups, tile_bytes) != cudaSuccess) return -2;
        if (rw.d_tiles != before) rw.tag = 0;
    }

    MmaParams P{};
    static long long *d_prof = nullptr;
    static const bool want_prof = getenv("MRB_MMA_PROF") != nullptr;
    if (want_prof && !d_prof) { cudaMalloc(&d_prof, 16 * 8 * 4096); cudaMemset(d_prof, 0, 16 * 8 * 4096); }
    P.prof = want_prof ? d_prof : nullptr;
    P.g_begin = g_begin; P.g_end = groups; P.y0 = y0; P.KB = KB; P.KS = KS; P.NWB = nwb; P.tile_bytes = tile_bytes;
    P.period = (int)ntiles; P.resident = resident ? 1 : 0;
    static const int fwd = getenv("MRB_MMA_FWD") && atoi(getenv("MRB_MMA_FWD")) == 0 ? 0 : 1;
    P.fwd = fwd;
    P.split = split ? 1 : 0;
    static const int n_issuers = getenv("MRB_MMA_ISSUERS") && atoi(getenv("MRB_MMA_ISSUERS")) == 1 ? 1 : 2;
    P.nissue = n_issuers;
    P.nch = (int)G.nch; P.hist = static_cast<const float *>(G.hist); P.H = split ? G.H / 2 : G.H;   // (G is the float view)
    const int64_t span = groups - g_begin;
    const int64_t cgroups = ceil_div(G.nch, kMmaRows);
    // time tiles: one CTA per SM (shared and tensor memory), so the grid should be whole waves of num_sms CTAs: take the
    // tile count that wastes the fewest SM-waves while keeping at least ~48 groups per tile (a CTA's start-up -- tensor
    // memory allocation, first loads -- costs about as much as 8 groups)
    int64_t tiles = 1;
    {
        const int64_t tmax = std::max<int64_t>(1, span / 48);
        double best = -1.0;
        for (int64_t t = 1; t <= std::min<int64_t>(tmax, 4096); ++t) {
            const int64_t ctas = t * cgroups, waves = ceil_div(ctas, (int64_t)p.num_sms);
            const double eff = (double)ctas / (double)(waves * p.num_sms) * (1.0 - 8.0 / (8.0 + (double)span / (double)t));
            if (eff > best + 1e-9) { best = eff; tiles = t; }
        }
    }
    if (ceil_div(span, tiles) > kMmaMaxGT) tiles = ceil_div(span, kMmaMaxGT);
    P.GT = (int)ceil_div(span, tiles);
    tiles = ceil_div(span, P.GT);

    CUtensorMap tmx, tmy;
    cuuint64_t dims[2] = {(cuuint64_t)G.n_in, (cuuint64_t)G.nch};
    cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 4};
    cuuint32_t box[2] = {kMmaBox, kMmaRows};
    cuuint32_t ones[2] = {1, 1};
    if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, ones,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        MRB_MMA_SKIP("x tensor map");
    cuuint64_t ydims[2] = {(cuuint64_t)(y0 + cnt), (cuuint64_t)G.nch};
    cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * 4};
    cuuint32_t ybox[2] = {GG, kMmaRows};
    if (p.encode(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G.y, ydims, ystrides, ybox, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        MRB_MMA_SKIP("y tensor map");
#undef MRB_MMA_SKIP
    if (tag == 0 || rw.tag != tag) {
        // pre-pass (launched only now: all host-side preparation is done, the two kernels go out back to back): one warp
        // per tile row (rows past the last output are zero) and one thread per group for its window start.  Skipped when
        // an earlier channel block of the same call already built this slice's tiles on this stream.
        rw.tag = tag;
        const int64_t nrows = ntiles * GG;
        const unsigned gb = (unsigned)std::max(ceil_div(nrows, 8), ceil_div(groups, 256));
        const int64_t outs = split ? (S.mode == 2 ? ntiles * (GG / 2) : cnt / 2) : (S.mode == 2 ? nrows : cnt);
        k_mma_tiles<GG><<<gb, 256, 0, st>>>((const float *)d_pfb, (const float *)d_dpfb, d_pnfb, P1, p.T, KB, S, split ? G.H / 2 : G.H,
                                            outs, nrows, groups, rw.d_tiles, rw.d_gstart);
        ++*launches;
        if (trace && cudaPeekAtLastError() != cudaSuccess)
            fprintf(stderr, "[mrb] k_mma_tiles failed: %s (mode %d cnt %lld groups %lld ntiles %lld KB %d)\n", cudaGetErrorString(cudaPeekAtLastError()),
                    S.mode, (long long)cnt, (long long)groups, (long long)ntiles, KB);
    }
    dim3 grid((unsigned)cgroups, (unsigned)tiles);
    if (split) k_mma_fir<GG, true><<<grid, kMmaThreads, fixed + nwb * tile_bytes, st>>>(tmx, tmy, rw.d_tiles, rw.d_gstart, P);
    else k_mma_fir<GG, false><<<grid, kMmaThreads, fixed + nwb * tile_bytes, st>>>(tmx, tmy, rw.d_tiles, rw.d_gstart, P);
    if (cudaPeekAtLastError() != cudaSuccess) {
        if (trace)
            fprintf(stderr, "[mrb] k_mma_fir failed: %s (mode %d cnt %lld y0 %lld n_in %lld nch %lld groups %lld tiles %lld GT %d KB %d KS %d NWB %d period %d resident %d span %lld H %lld)\n",
                    cudaGetErrorString(cudaPeekAtLastError()), S.mode, (long long)cnt, (long long)y0, (long long)G.n_in, (long long)G.nch, (long long)groups,
                    (long long)tiles, P.GT, KB, KS, nwb, P.period, P.resident, (long long)max_group_span, (long long)G.H);
        return -2;
    }
    if (want_prof && (int64_t)cgroups * tiles <= 4096) {
        static int shown = 0;
        if (shown++ < 3) {
            std::vector<long long> hp((size_t)cgroups * tiles * 16);
            cudaStreamSynchronize(st);
            cudaMemcpy(hp.data(), d_prof, hp.size() * 8, cudaMemcpyDeviceToHost);
            double s[16] = {};
            for (size_t i = 0; i < hp.size(); ++i) s[i % 16] += (double)hp[i] / (double)(cgroups * tiles);
            fprintf(stderr, "[mrb] mma prof (mean cycles per CTA, %lld CTAs, %.0f groups, %.0f boxes): xload wait-empty %.0f | wload wait-empty %.0f | "
                            "mma wait a_full %.0f w_full %.0f d_empty %.0f fence %.0f issue %.0f total %.0f | conv wait x_full %.0f a_empty %.0f total %.0f | "
                            "epi wait d_full %.0f tma-read %.0f total %.0f\n",
                    (long long)(cgroups * tiles), s[9], s[10], s[0], s[1], s[2], s[3], s[4], s[15], s[11], s[14], s[5], s[6], s[13], s[7], s[8], s[12]);
        }
    }
    *name = S.cplx == 2 ? (resident ? "mma_c64_split_resident" : "mma_c64_split") : S.cplx ? (resident ? "mma_c64_g32_resident" : "mma_c64_g32") : (resident ? "mma_f32_g32_resident" : "mma_f32_g32");
    ++*launches;
    return k_begin;
}

}  // namespace mrb
