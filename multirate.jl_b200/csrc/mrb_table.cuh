// mrb_table.cuh -- fast path for the arbitrary-rate kernels on real samples: FIRArbitrary (src/Filters.jl:693-742)
// and FIRFarrow (src/Filters.jl:795-836); BASELINE configs[3].
//
// Both kernels compute, per output k, one dot product of a per-output tap row with the window of x that the exact
// host replay of the phase recurrence (mrb_seq.h, src/Filters.jl:663-673 / 780-786) assigns to it:
//   arbitrary:  taps_k[i] = pfb[i, phi_k] + alpha_k * dpfb[i, phi_k]     (the reference's tapsforphase, :681-686;
//               filt! itself blends the two dot products, :730 -- algebraically the same, rounding differs by
//               ~1e-7 relative, inside the stated tolerance)
//   farrow:     taps_k[i] = polyval(pnfb[i], phase_k)                     (:789-791)
// The tap rows are data independent and shared by every channel, so a pre-pass builds them once per chunk
// (k_table_rows), already shifted so that the dot product can start at a 16-byte aligned sample:
//   row_k[j] = taps_k[j - d_k],  d_k = (window start of output k) mod (4 floats | 2 doubles),  zero elsewhere.
// The main kernel (k_table_fir) is then one fixed-length dot product per output and channel:
//  * lane = channel; a CTA is 32 channels x 4 warps, a step is 32 consecutive outputs (8 per warp);
//  * samples arrive by TMA in [32 ch][128-byte] boxes (SWIZZLE_128B) in a ring; a lane reads its window with
//    conflict-free LDS.128; the tap row is read with warp-uniform 128-bit loads (one L1 transaction each);
//  * a step's 32 outputs per channel are staged in shared memory and leave with one TMA store.
// Outputs whose window touches the history are computed by k_generic.
#pragma once
#include <cstdio>

#include "mrb_tiled.cuh"

namespace mrb {

constexpr int kTabRows = 32;            // channels per CTA
constexpr int kTabStep = 32;            // outputs per CTA step
constexpr int kTabWarps = 4;
constexpr int kTabGroup = 8;            // consecutive outputs that share one register window (one warp's share of a step)
constexpr int kTabNB2 = 12;             // ring boxes of the two-channels-per-lane variant (64-row boxes of 8 KB)
constexpr int kTabTB2 = 24;             // its tap block (Float64 samples): 24, or 22 when the rows fit

// sample kinds the kernel is instantiated for
enum { TAB_F32 = 0, TAB_F64 = 1, TAB_C64 = 2 };
template <int K> struct TabCfg;
template <> struct TabCfg<TAB_F32> {
    using Tap = float;                  // tap-row element
    static constexpr int ES = 4;        // bytes per sample
    static constexpr int A = 4;         // samples per 16 bytes
    static constexpr int TB = 96;       // row elements (window samples) per block
    static constexpr int BOXE = 32;     // samples per box row (128 B)
    static constexpr int NB = 10;       // ring boxes
};
template <> struct TabCfg<TAB_F64> {
    using Tap = double;
    static constexpr int ES = 8;
    static constexpr int A = 2;
    static constexpr int TB = 48;
    static constexpr int BOXE = 16;
    static constexpr int NB = 14;
};
template <> struct TabCfg<TAB_C64> {    // complex64 samples x float32 taps: two FMAs per tap
    using Tap = float;
    static constexpr int ES = 8;
    static constexpr int A = 2;
    static constexpr int TB = 48;
    static constexpr int BOXE = 16;
    static constexpr int NB = 14;
};

// ---------------------------------------------------------------------------------------------------------
// pre-pass: tap rows + aligned window starts for outputs [0, nout) of a schedule slice
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256)
k_table_rows(const R *__restrict__ pfb, const R *__restrict__ dpfb, const double *__restrict__ pnfb, int P1, int T,
             int rowlen, int farrow, int tap_is_f32, const int64_t *__restrict__ sn, const int32_t *__restrict__ sphi,
             const double *__restrict__ sa, int64_t H, int64_t nout, R *__restrict__ rows, int32_t *__restrict__ astart,
             int A, int64_t iL = 0, int64_t iM = 0, int64_t ip0 = 0, int64_t id0m1 = 0) {
    // one warp per output row (8 rows per block): the row's schedule entries are read once, no index division
    const int64_t k = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= nout) return;
    const int lane = threadIdx.x & 31;
    // the kTabGroup outputs of a group read ONE register window that starts at the aligned window start of the
    // group's first output; every row of the group is shifted to its own place inside that window
    // sn == nullptr: an integer schedule in closed form (src/Filters.jl:558-569: n_k = d + floor((p + k M) / L), branch
    // (p + k M) mod L; no blend) -- rational / interpolator / standard filters on Float64 samples
    const bool closed = sn == nullptr;
    const int64_t kg0 = k / kTabGroup * kTabGroup;
    const int64_t xs = (closed ? id0m1 + (ip0 + k * iM) / iL : sn[k]) - H;       // x index of the window start (may be < 0: head)
    const int64_t xg = (closed ? id0m1 + (ip0 + kg0 * iM) / iL : sn[kg0]) - H;
    const int64_t al = xg >= 0 ? xg / A * A : -((-xg + A - 1) / A) * A;
    const int d = (int)(xs - al);
    if (lane == 0) astart[k] = (int32_t)al;
    const double ph = closed ? 0.0 : sa[k];                          // farrow: phase; arbitrary: alpha
    const int64_t obase = farrow ? 0 : (closed ? (ip0 + k * iM) % iL : (int64_t)sphi[k]) * T;
    for (int j = lane; j < rowlen; j += 32) {
        const int i = j - d;
        R v = R(0);
        if (i >= 0 && i < T) {
            if (farrow) {
                // currentTaps[i] = polyval(pnfb[i], phase): Horner highest order first in Float64 with separately
                // rounded multiply and add, rounded to the tap type (src/Filters.jl:789-791)
                const double *c = pnfb + (int64_t)i * P1;
                double a = c[P1 - 1];
                for (int p = P1 - 2; p >= 0; --p) a = __dadd_rn(__dmul_rn(a, ph), c[p]);
                if (tap_is_f32) a = (double)(float)a;
                v = (R)a;
            } else {
                v = closed ? pfb[obase + i] : (R)((double)pfb[obase + i] + ph * (double)dpfb[obase + i]);
            }
        }
        rows[k * rowlen + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------
struct alignas(16) TabParams {
    long long k_begin, N;      // outputs [k_begin, N) of the slice (slice-relative; k_begin multiple of 32)
    long long y0;              // output index of slice output 0 in y (multiple of 4)
    int KT;                    // outputs per tile (multiple of 32)
    int nblk;                  // tap blocks per row
    int rowlen;                // nblk * TB
    int pad0, pad1, pad2;
    unsigned win[4][160];      // win[c][i] = W((c + i) mod 8 NB): chunk bits | ring slot offset of 16-byte chunk u
};

// TB = row elements (window samples) per tap block: TabCfg<K>::TB, or the next smaller size when the rows fit it
// (every element of a row is a tap load and an FMA for all channels, zero padding included)
// CH = channels per lane (the CTA is 32 CH channels wide).  CH = 2 (Float64): every broadcast tap load feeds the FMAs of two
// channels -- the kernel is bound by shared-memory loads (90 % of the LSU at CH = 1), not by the FP64 pipe -- with tap
// blocks of TB = 22/24 so that the two windows stay in registers (2 x 24 doubles) and ONE accumulator per output.
// WPG = warps per window group of 8 outputs.  Measured on C4 Float64: two warps per group (8 warps per SM, 4 outputs each)
// run 6 % SLOWER than one (64.0 against 68.0 Gout/s) -- the kernel is not short of warps, so WPG stays 1.
constexpr int kTabWPG = 1;
// DM = 4 or 8 (Float64, CH = 2): the FP64 tensor-core variant with DM warps.  A group of 8 outputs x 8 channels is one
// accumulator tile of mma.sync.m8n8k4.f64: A = the window [8 channels x 4 samples] (one LDS.64 per lane from the swizzled
// ring, 2 wavefronts), B = the group's shifted tap rows [4 window positions x 8 outputs] (one LDS.64 per lane, the row
// pitch = 4 mod 8 doubles keeps it at 2 wavefronts), reused by every channel octet of the warp.  Against DFMA this is one
// instruction and 16 loaded bytes per 256 FMAs instead of per 32 / 64: the kernel leaves the LSU wall (DFMA and DMMA peak at
// the same 37 TFLOP/s on this part, profiles/r2_dfma_probe.txt).  TB = 4 (one k-slice) in this variant.
template <int K, int TB, int CH, int DM = 0>
__global__ void __launch_bounds__(DM == 8 ? 288 : DM ? DM * 32 : 128 * kTabWPG, CH == 1 ? 3 : 1)
k_table_fir(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
            const typename TabCfg<K>::Tap *__restrict__ rows, const int32_t *__restrict__ astart,
            const __grid_constant__ TabParams P) {
    using C = TabCfg<K>;
    using R = typename C::Tap;
    constexpr int A = C::A, NQ = TB / A, ES = C::ES;
    constexpr int NB = CH == 1 ? C::NB : kTabNB2;                     // ring boxes
    constexpr int ROWS = kTabRows * CH;                               // channels per CTA
    static_assert(CH == 1 || K == TAB_F64, "two channels per lane: Float64 only");
    static_assert(TB % 2 == 0 && TB <= C::TB, "tap rows are read four (two) at a time");
    static_assert(DM == 0 || (CH == 2 && TB == 4 && (DM == 4 || DM == 8)), "tensor-core variant: 64-channel CTAs, k-slices of 4");
    constexpr int BOX_BYTES = ROWS * 128;
    constexpr int WPG = kTabWPG;                                     // warps per window group
    constexpr int OPW = kTabStep / kTabWarps / WPG;                  // outputs per warp per step
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *out_buf = smem + NB * BOX_BYTES;                  // [32 ch][32 outputs], 128-byte swizzle atoms
    // the tap rows and aligned window starts of a step (32 outputs), double buffered, fetched by bulk copies a
    // step ahead: the taps are then read with warp-uniform LDS.128 instead of L2-latency global loads
    const int row_bytes = kTabStep * P.rowlen * (int)sizeof(R);      // R = tap type
    // DB: two staging buffers and ONE CTA barrier per step (the tensor-core variant, one CTA per SM: a warp that waits at a
    // barrier is a quarter of the SM idle) -- thread 0 makes sure the previous step's store has left its buffer BEFORE the
    // barrier, issues this step's store after it and does not wait for it
    constexpr bool DB = DM != 0;
    // PW (DM = 8): a ninth warp takes thread 0's per-step duties (output store, ring refill, next tap rows) and the CTA barrier
    // goes: consumers arrive on done[s & 1] when their part of the step is staged, the producer signals free[b] when the store
    // that read staging buffer b has finished -- the consumer warps run from step to step without meeting
    constexpr bool PW = DM == 8;
    constexpr int STG_BYTES = ROWS * kTabStep * ES;
    unsigned char *rows_s = out_buf + (DB ? 2 : 1) * STG_BYTES;
    int *ast_s = reinterpret_cast<int *>(rows_s + 2 * row_bytes);    // [2][32]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ast_s + 2 * kTabStep);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ch0 = blockIdx.x * ROWS;
    const uint32_t in_base = smem_u32(smem), obase = smem_u32(out_buf), bar_base = smem_u32(bars);
    const uint32_t rbar_base = bar_base + 8 * NB;                    // two mbarriers for the staged rows
    const uint32_t done_base = rbar_base + 16, free_base = done_base + 16;   // PW: two "step staged", two "staging free"
    const uint32_t rowpart = ((uint32_t)lane * 128u) ^ (((uint32_t)lane & 7u) << 4);   // SWIZZLE_128B
    auto stage_rows = [&](long long kstep, int slot) {               // thread 0 only
        const uint32_t bar = rbar_base + 8 * slot;
        mbar_expect_tx(bar, (uint32_t)row_bytes + kTabStep * 4);
        bulk_load(smem_u32(rows_s) + (uint32_t)(slot * row_bytes), rows + kstep * P.rowlen, (uint32_t)row_bytes, bar);
        bulk_load(smem_u32(ast_s) + (uint32_t)(slot * kTabStep * 4), astart + kstep, kTabStep * 4, bar);
    };

    const long long k0 = P.k_begin + (long long)blockIdx.y * P.KT;   // first output of the tile
    const int ntile = (int)min((long long)P.KT, P.N - k0);
    const int nsteps = (ntile + kTabStep - 1) / kTabStep;
    const long long klast = k0 + ntile - 1;
    const int xbase = astart[k0] / C::BOXE * C::BOXE;                // element index of box 0 (astart >= 0 here)
    const int jlast = (astart[klast] + P.rowlen - 1 - xbase) / C::BOXE;   // newest box the tile reads

    if (tid == 0) {
        if (in_base & 1023u) __trap();
        for (int i = 0; i < NB + 2; ++i) mbar_init(bar_base + 8 * i, 1);
        if constexpr (PW) {
            mbar_init(done_base, DM); mbar_init(done_base + 8, DM);
            mbar_init(free_base, 1); mbar_init(free_base + 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
        stage_rows(k0, 0);
        if (nsteps > 1) stage_rows(k0 + kTabStep, 1);
        for (int jj = 0; jj < NB && jj <= jlast; ++jj) {
            mbar_expect_tx(bar_base + 8 * jj, BOX_BYTES);
            tma_load_2d(in_base + (uint32_t)(jj * BOX_BYTES), &tmx, xbase + jj * C::BOXE, ch0, bar_base + 8 * jj);
        }
    }
    __syncthreads();

    int j_issued = min(NB, jlast + 1), i_slot = j_issued % NB;
    int j_waited = 0, w_slot = 0;
    uint32_t w_par = 0;

    if constexpr (PW) {
        if (warp == DM) {                                            // ---- producer warp
            if (lane == 0) {
                for (int s = 0; s < nsteps; ++s) {
                    const long long ks = k0 + (long long)s * kTabStep;
                    const int rslot = s & 1;
                    mbar_wait(done_base + 8 * rslot, (uint32_t)((s >> 1) & 1));       // every warp has staged step s
                    constexpr int NST = kTabStep * ES / 128;
                    for (int b = 0; b < NST; ++b)
                        tma_store_2d(&tmy, (int)(P.y0 + ks) + b * C::BOXE, ch0, obase + (uint32_t)(rslot * STG_BYTES + b * ROWS * 128));
                    tma_commit();
                    int jtarget = jlast;
                    if (s + 1 < nsteps) {                            // the next step's first window start is already in shared memory
                        mbar_wait(rbar_base + 8 * (rslot ^ 1), (uint32_t)(((s + 1) >> 1) & 1));
                        jtarget = min((ast_s[(rslot ^ 1) * kTabStep] - xbase) / C::BOXE + NB - 1, jlast);
                    }
                    for (int jj = j_issued; jj <= jtarget; ++jj) {
                        const uint32_t bar = bar_base + 8 * i_slot;
                        mbar_expect_tx(bar, BOX_BYTES);
                        tma_load_2d(in_base + (uint32_t)(i_slot * BOX_BYTES), &tmx, xbase + jj * C::BOXE, ch0, bar);
                        if (++i_slot == NB) i_slot = 0;
                    }
                    if (jtarget >= j_issued) j_issued = jtarget + 1;
                    if (s + 2 < nsteps) stage_rows(ks + 2 * kTabStep, rslot);
                    tma_wait_read<1>();                              // the store of step s - 1 has left its buffer
                    if (s >= 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(free_base + 8 * (uint32_t)(rslot ^ 1)) : "memory");
                }
                tma_wait_read<0>();
            }
            return;
        }
    }

    for (int s = 0; s < nsteps; ++s) {
        const long long ks = k0 + (long long)s * kTabStep;
        const int rslot = s & 1;
        mbar_wait(rbar_base + 8 * rslot, (uint32_t)((s >> 1) & 1));  // this step's tap rows are in shared memory
        const R *rows_step = reinterpret_cast<const R *>(rows_s + rslot * row_bytes);
        const int *ast_step = ast_s + rslot * kTabStep;
        if constexpr (DM != 0) {
            // ---- FP64 tensor cores: warp = (group of 8 outputs, DM == 8: half of the channels)
            constexpr int CB = DM == 4 ? 8 : 4;                      // channel octets per warp
            const int grp = warp & 3, half = DM == 4 ? 0 : warp >> 2;
            const int r = lane >> 2, kq = lane & 3;                  // fragment row (channel / output), k index
            // fragment row r <-> channel rr of the octet: a 64-bit shared load is served per half warp (fragment rows 0..3),
            // whose four 32-byte reads must fall into four different chunk pairs of the 128-byte swizzle: rows 0, 2, 4, 6
            // (rows 0..3 share two pairs: every A load cost 4 wavefronts instead of 2)
            const int rr = ((r & 3) << 1) | (r >> 2);
            const long long kg = ks + grp * kTabGroup;
            double acc[CB][2];
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[c][0] = acc[c][1] = 0.0;
            if (kg <= klast) {
                const int a0 = ast_step[grp * kTabGroup] - xbase;    // tile-relative aligned window start (even)
                const int need = (a0 + P.rowlen - 1) / C::BOXE;
                for (; j_waited <= need; ++j_waited) {
                    mbar_wait(bar_base + 8 * w_slot, w_par);
                    if (++w_slot == NB) { w_slot = 0; w_par ^= 1u; }
                }
                // B fragment: lane (k = kq, n = r) reads row r of the group at window position 4 kk + kq
                const uint32_t bp = smem_u32(rows_step) + (uint32_t)(((grp * kTabGroup + r) * P.rowlen + kq) * 8);
                // A fragment: lane (m = r, k = kq) reads channel 8 c + r at sample a0 + 4 kk + kq: ring chunk u, half kq & 1
                int u = (a0 >> 1) % (8 * NB) + (kq >> 1);
                const uint32_t abase = in_base + (uint32_t)(half * (CB * 8 * 128) + rr * 128 + (kq & 1) * 8);
                // software pipeline: the fragments of slice kk + 1 are in flight while the tensor pipe works on slice kk
                auto load_slice = [&](int kk, double &b, double (&a)[CB]) {
                    if (u >= 8 * NB) u -= 8 * NB;
                    const uint32_t ad = abase + (uint32_t)((((u & 7) ^ rr) << 4) + (u >> 3) * BOX_BYTES);
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b) : "r"(bp + (uint32_t)(kk * 32)) : "memory");
#pragma unroll
                    for (int c = 0; c < CB; ++c)
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[c]) : "r"(ad + (uint32_t)(c * 8 * 128)) : "memory");
                    u += 2;
                };
                double b0, a0f[CB], b1, a1f[CB];
                load_slice(0, b0, a0f);
                for (int kk = 0; kk < P.nblk; kk += 2) {             // nblk is odd: the second half of the last pair is skipped
                    const bool more = kk + 1 < P.nblk;
                    if (more) load_slice(kk + 1, b1, a1f);
#pragma unroll
                    for (int c = 0; c < CB; ++c)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(acc[c][0]), "+d"(acc[c][1]) : "d"(a0f[c]), "d"(b0));
                    if (!more) break;
                    if (kk + 2 < P.nblk) load_slice(kk + 2, b0, a0f);
#pragma unroll
                    for (int c = 0; c < CB; ++c)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(acc[c][0]), "+d"(acc[c][1]) : "d"(a1f[c]), "d"(b1));
                }
            }
            // stage: lane holds channel 8 c + rr, outputs 2 kq and 2 kq + 1 of the group
            if constexpr (PW) {
                if (s >= 2) mbar_wait(free_base + 8 * (uint32_t)(s & 1), (uint32_t)(((s >> 1) - 1) & 1));   // store s - 2 is done
            }
            const uint32_t byte = (uint32_t)(grp * kTabGroup + 2 * kq) * 8u;
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                const uint32_t row = (uint32_t)(half * CB * 8 + c * 8 + rr);
                const uint32_t ad = obase + (uint32_t)((s & 1) * STG_BYTES) + (byte >> 7) * (ROWS * 128u) + row * 128u +
                                    ((((byte >> 4) & 7u) ^ (row & 7u)) << 4);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ad), "d"(acc[c][0]), "d"(acc[c][1]) : "memory");
            }
        } else {
            static_assert(OPW * WPG == kTabGroup, "the warps of a group share one window");
            const int grp = warp / WPG, sub = warp % WPG;            // window group of this warp, its part of the group
            const long long kg = ks + grp * kTabGroup + sub * OPW;   // first output this warp computes
            // f32 / f64: two partial sums per output; c64: real and imaginary part; CH = 2: one sum per output and channel
            R acc[OPW][2];
#pragma unroll
            for (int o = 0; o < OPW; ++o) acc[o][0] = acc[o][1] = R(0);
            if (kg <= klast) {
                const int a0 = ast_step[grp * kTabGroup] - xbase;    // tile-relative aligned window start (samples)
                const int need = (a0 + P.rowlen - 1) / C::BOXE;
                for (; j_waited <= need; ++j_waited) {
                    mbar_wait(bar_base + 8 * w_slot, w_par);
                    if (++w_slot == NB) { w_slot = 0; w_par ^= 1u; }
                }
                const R *rowg = rows_step + (grp * kTabGroup + sub * OPW) * P.rowlen;
                for (int bb = 0; bb < P.nblk; ++bb) {
                    const int p = ((a0 + bb * TB) / A) % (8 * NB);   // ring position in 16-byte chunks
                    constexpr int WR = K == TAB_C64 ? 2 * TB : TB;   // window registers of type R (per channel)
                    R w[CH][WR];
                    if constexpr (CH == 1) {
                        const unsigned *wt = P.win[p & 3] + (p & ~3);
#pragma unroll
                        for (int q = 0; q < NQ; q += 4) {
                            const uint4 w4 = *reinterpret_cast<const uint4 *>(wt + q);
                            const unsigned ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (q + e >= NQ) break;
                                const uint32_t ad = in_base + (rowpart ^ ww[e]);
                                if constexpr (K == TAB_F64) {
                                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                                                 : "=d"(w[0][2 * (q + e)]), "=d"(w[0][2 * (q + e) + 1]) : "r"(ad) : "memory");
                                } else {                             // four floats: 4 real samples or 2 complex ones
                                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                                 : "=f"(w[0][4 * (q + e)]), "=f"(w[0][4 * (q + e) + 1]), "=f"(w[0][4 * (q + e) + 2]),
                                                   "=f"(w[0][4 * (q + e) + 3]) : "r"(ad) : "memory");
                                }
                            }
                        }
                    } else {
                        // ring position -> address word computed here (the CH = 1 table in the parameter block is laid out for
                        // 32-row boxes): chunk u of the ring lives in box u >> 3, 16-byte chunk u & 7
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            int u = p + q;
                            if (u >= 8 * NB) u -= 8 * NB;
                            const uint32_t word = (uint32_t)((u & 7) << 4) | (uint32_t)((u >> 3) * BOX_BYTES);
#pragma unroll
                            for (int c = 0; c < CH; ++c) {
                                const uint32_t ad = in_base + ((rowpart + (uint32_t)(c * kTabRows * 128)) ^ word);
                                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w[c][2 * q]), "=d"(w[c][2 * q + 1]) : "r"(ad) : "memory");
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < OPW; ++o) {
                        const R *tr = rowg + o * P.rowlen + bb * TB;
                        if constexpr (CH == 2) {
#pragma unroll
                            for (int q = 0; q < TB / 2; ++q) {
                                const double2 t = reinterpret_cast<const double2 *>(tr)[q];
                                acc[o][0] = fma(t.x, w[0][2 * q], acc[o][0]);
                                acc[o][1] = fma(t.x, w[1][2 * q], acc[o][1]);
                                acc[o][0] = fma(t.y, w[0][2 * q + 1], acc[o][0]);
                                acc[o][1] = fma(t.y, w[1][2 * q + 1], acc[o][1]);
                            }
                        } else if constexpr (K == TAB_F32) {
                            // the two partial sums of an output are one packed pair: (even taps, odd taps) x (even
                            // samples, odd samples) is one FFMA2 on two aligned register pairs -- half the issue slots
                            unsigned long long a2;
                            asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(acc[o][0]), "f"(acc[o][1]));
#pragma unroll
                            for (int q = 0; q < TB / 4; ++q) {
                                const float4 t = reinterpret_cast<const float4 *>(tr)[q];
                                unsigned long long t01, t23, w01, w23;
                                asm("mov.b64 %0, {%1, %2};" : "=l"(t01) : "f"(t.x), "f"(t.y));
                                asm("mov.b64 %0, {%1, %2};" : "=l"(t23) : "f"(t.z), "f"(t.w));
                                asm("mov.b64 %0, {%1, %2};" : "=l"(w01) : "f"(w[0][4 * q]), "f"(w[0][4 * q + 1]));
                                asm("mov.b64 %0, {%1, %2};" : "=l"(w23) : "f"(w[0][4 * q + 2]), "f"(w[0][4 * q + 3]));
                                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2) : "l"(t01), "l"(w01));
                                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2) : "l"(t23), "l"(w23));
                            }
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[o][0]), "=f"(acc[o][1]) : "l"(a2));
                        } else if constexpr (K == TAB_F64) {
#pragma unroll
                            for (int q = 0; q < TB / 2; ++q) {
                                const double2 t = reinterpret_cast<const double2 *>(tr)[q];
                                acc[o][0] = fma(t.x, w[0][2 * q], acc[o][0]);
                                acc[o][1] = fma(t.y, w[0][2 * q + 1], acc[o][1]);
                            }
                        } else {                                     // complex sample i = (w[2i], w[2i+1]), real tap
                            // one FFMA2 per tap: (re, im) += t * (re, im), the tap a scalar register operand
                            unsigned long long a2;
                            asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(acc[o][0]), "f"(acc[o][1]));
#pragma unroll
                            for (int q = 0; q < TB / 4; ++q) {
                                const float4 t = reinterpret_cast<const float4 *>(tr)[q];
                                const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    unsigned long long xs;
                                    asm("mov.b64 %0, {%1, %2};" : "=l"(xs) : "f"(w[0][8 * q + 2 * e]), "f"(w[0][8 * q + 2 * e + 1]));
                                    cfma(a2, tt[e], xs);
                                }
                            }
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[o][0]), "=f"(acc[o][1]) : "l"(a2));
                        }
                    }
                }
            }
            // stage: row = channel, column = output within the step; 128-byte swizzle atoms of 32/16 outputs
#pragma unroll
            for (int o = 0; o < OPW; ++o) {
                const int col = grp * kTabGroup + sub * OPW + o;
                const uint32_t byte = (uint32_t)col * ES;
                const uint32_t ad = obase + (byte >> 7) * (ROWS * 128u) + (rowpart ^ (((byte >> 4) & 7u) << 4)) + (byte & 15u);
                if constexpr (CH == 2) {
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad), "d"(acc[o][0]) : "memory");
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad + kTabRows * 128u), "d"(acc[o][1]) : "memory");
                } else if constexpr (K == TAB_F32) {
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad), "f"(acc[o][0] + acc[o][1]) : "memory");
                } else if constexpr (K == TAB_F64) {
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad), "d"(acc[o][0] + acc[o][1]) : "memory");
                } else {
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ad), "f"(acc[o][0]), "f"(acc[o][1]) : "memory");
                }
            }
        }

        // ---- the step's outputs leave; boxes before the next step's first window are refilled
        fence_async_smem();
        if constexpr (PW) {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(done_base + 8 * (uint32_t)(s & 1)) : "memory");
            continue;
        }
        if constexpr (DB) {
            if (tid == 0) tma_wait_read<0>();              // the previous step's store has left the other staging buffer
            __syncthreads();
            if (tid == 0) {
                constexpr int NST = kTabStep * ES / 128;
                for (int b = 0; b < NST; ++b)
                    tma_store_2d(&tmy, (int)(P.y0 + ks) + b * C::BOXE, ch0, obase + (uint32_t)((s & 1) * STG_BYTES + b * ROWS * 128));
                tma_commit();
                int jtarget = jlast;
                if (s + 1 < nsteps) {                      // the next step's first window start is already in shared memory
                    mbar_wait(rbar_base + 8 * (rslot ^ 1), (uint32_t)(((s + 1) >> 1) & 1));
                    jtarget = min((ast_s[(rslot ^ 1) * kTabStep] - xbase) / C::BOXE + NB - 1, jlast);
                }
                for (int jj = j_issued; jj <= jtarget; ++jj) {
                    const uint32_t bar = bar_base + 8 * i_slot;
                    mbar_expect_tx(bar, BOX_BYTES);
                    tma_load_2d(in_base + (uint32_t)(i_slot * BOX_BYTES), &tmx, xbase + jj * C::BOXE, ch0, bar);
                    if (++i_slot == NB) i_slot = 0;
                }
                if (jtarget >= j_issued) j_issued = jtarget + 1;
                if (s + 2 < nsteps) stage_rows(ks + 2 * kTabStep, rslot);
            }
            continue;
        }
        __syncthreads();
        const long long kn = min(ks + kTabStep, klast);
        const int jdead = (astart[kn] - xbase) / C::BOXE;            // oldest box the next step reads
        const int jtarget = min(jdead + NB - 1, jlast);
        if (tid == 0) {
            constexpr int NST = kTabStep * ES / 128;                 // 128-byte-wide sub-boxes per step
            for (int b = 0; b < NST; ++b)
                tma_store_2d(&tmy, (int)(P.y0 + ks) + b * C::BOXE, ch0, obase + (uint32_t)(b * ROWS * 128));
            tma_commit();
            int sl = i_slot;
            for (int jj = j_issued; jj <= jtarget; ++jj) {
                const uint32_t bar = bar_base + 8 * sl;
                mbar_expect_tx(bar, BOX_BYTES);
                tma_load_2d(in_base + (uint32_t)(sl * BOX_BYTES), &tmx, xbase + jj * C::BOXE, ch0, bar);
                if (++sl == NB) sl = 0;
            }
            // the rows of step s+2 go where this step's rows were (every warp is past them: barrier above)
            if (s + 2 < nsteps) stage_rows(ks + 2 * kTabStep, rslot);
            tma_wait_read<0>();                            // the staging buffer is rewritten in the next step
        }
        if (jtarget >= j_issued) {
            i_slot = (i_slot + (jtarget + 1 - j_issued)) % NB;
            j_issued = jtarget + 1;
        }
        __syncthreads();
    }
    if (DB && !PW && tid == 0) tma_wait_read<0>();        // the last store still reads its staging buffer (PW: the producer waits)
    for (; j_waited < j_issued; ++j_waited) {             // every issued load must have landed before exit (DB: thread 0 knows)
        mbar_wait(bar_base + 8 * w_slot, w_par);
        if (++w_slot == NB) { w_slot = 0; w_par ^= 1u; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// Device buffers one schedule slice's tap rows live in.  They are WRITTEN by every call (k_table_rows), so the
// handle keeps one set per pipeline stream (mrb_api.cu TableCtx), not one per plan.
struct TabRows {
    void *d_rows = nullptr;            // Tap[slice outputs][rowlen]
    int32_t *d_astart = nullptr;
    int64_t cap = 0;                   // outputs the two buffers hold
    int64_t row_bytes = 0;             // rowlen * tap bytes the buffers were sized for
    uint64_t tag = 0;                  // (call serial, slice) the rows were built for: 0 = none
};

static inline void tabrows_release(TabRows &r) {
    cudaFree(r.d_rows); cudaFree(r.d_astart);
    r = TabRows{};
}

struct TabPlan {
    bool ok = false;
    int K = TAB_F32;                   // sample kind the plan was built for
    int es = 4, ts = 4, A = 4, TB = 96, NB = 10;   // sample bytes, tap bytes, samples per 16 B, block, ring boxes
    int T = 0, nblk = 0, rowlen = 0;
    int ch = 1;                        // channels per lane (2: Float64, the LSU-bound case)
    int dm = 0;                        // warps of the FP64 tensor-core variant (0: CUDA cores)
    TabParams *hp = nullptr;
    PFN_encodeTiled encode = nullptr;
    int num_sms = 148;
};

static inline void table_release(TabPlan &p) {
    delete p.hp;
    p.hp = nullptr;
    p.ok = false;
}

static inline int table_smem(const TabPlan &p) {
    return p.NB * p.ch * kTabRows * 128 + (p.dm ? 2 : 1) * p.ch * kTabRows * kTabStep * p.es + 2 * kTabStep * p.rowlen * p.ts + 2 * kTabStep * 4 +
           8 * (p.NB + 6);
}

// kind/tx/ty are the mrb.h enums (4 arbitrary, 5 farrow ; 0 = float32, 1 = float64, 2 = complex64)
static inline int32_t table_prepare(TabPlan &p, int kind, int tx, int ty, int64_t T, double rate, const cudaDeviceProp &prop) {
    p.ok = false;
    // arbitrary / farrow; and -- Float64 samples only, on the FP64 tensor-core variant -- the integer kinds whose windows fit the
    // ring (standard, interpolator, rational, gentle decimators: every Float64 call used to land on k_stream)
    const bool int_kind = kind >= 0 && kind <= 3;
    if (!(kind == 4 || kind == 5 || (int_kind && tx == 1 && ty == 1))) return 0;
    if (tx != ty || !(tx == 0 || tx == 1 || tx == 2)) return 0;         // float32, float64, complex64; no promotion
    p.K = tx == 0 ? TAB_F32 : tx == 1 ? TAB_F64 : TAB_C64;
    if (p.K == TAB_F32) { p.es = 4; p.ts = 4; p.A = 4; p.TB = TabCfg<TAB_F32>::TB; p.NB = TabCfg<TAB_F32>::NB; }
    else if (p.K == TAB_F64) { p.es = 8; p.ts = 8; p.A = 2; p.TB = TabCfg<TAB_F64>::TB; p.NB = TabCfg<TAB_F64>::NB; }
    else { p.es = 8; p.ts = 4; p.A = 2; p.TB = TabCfg<TAB_C64>::TB; p.NB = TabCfg<TAB_C64>::NB; }
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int32_t)(e ? e : cudaErrorUnknown);
    p.encode = (PFN_encodeTiled)fn;
    p.num_sms = prop.multiProcessorCount;
    p.T = (int)T;
    // a group's rows are shifted by up to (A-1) + (window start of its last output - that of its first)
    const int64_t gspan = rate > 0.0 ? (int64_t)std::ceil((kTabGroup - 1) / rate) + 1 : (int64_t)1 << 20;
    static const bool no2 = getenv("MRB_TABLE_CH1") != nullptr;
    p.ch = (p.K == TAB_F64 && !no2) ? 2 : 1;
    p.dm = 0;
    if (p.ch == 2) {
        // FP64 tensor cores (mma.sync m8n8k4): k-slices of 4, row pitch = 4 mod 8 doubles (conflict-free fragment loads)
        static const char *nd = getenv("MRB_NO_DMMA"), *dw = getenv("MRB_DMMA_WARPS");
        int64_t rl = T + p.A - 1 + gspan;
        rl = (rl + 3) / 8 * 8 + 4;
        TabPlan q = p;
        q.dm = dw && atoi(dw) == 4 ? 4 : 8; q.TB = 4; q.NB = kTabNB2; q.rowlen = (int)rl; q.nblk = (int)(rl / 4);
        if (!nd && rl <= 256 && table_smem(q) <= (int)prop.sharedMemPerBlockOptin) p = q;
    }
    if (int_kind && p.dm == 0) return 0;
    if (p.dm == 0) {
    if (p.ch == 2) { p.TB = kTabTB2; p.NB = kTabNB2; }
    p.nblk = (int)ceil_div(T + p.A - 1 + gspan, p.TB);
    if (p.nblk > (p.ch == 2 ? 4 : 2)) {                                  // taps too long / rate too low for a shared window
        if (p.ch == 2) {                                                 // (try the one-channel shape)
            p.ch = 1; p.TB = TabCfg<TAB_F64>::TB; p.NB = TabCfg<TAB_F64>::NB;
            p.nblk = (int)ceil_div(T + p.A - 1 + gspan, p.TB);
        }
        if (p.nblk > 2) return 0;
    }
    {   // the next smaller block (11 instead of 12 16-byte chunks per 24 taps) when the rows still fit the same number of blocks
        const int tbr = p.TB / 12 * 11;
        if (ceil_div(T + p.A - 1 + gspan, (int64_t)tbr) == p.nblk) p.TB = tbr;
    }
    }
    p.rowlen = p.nblk * p.TB;
    p.hp = new TabParams();
    memset(p.hp, 0, sizeof(TabParams));
    p.hp->nblk = p.nblk; p.hp->rowlen = p.rowlen;
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 160; ++i) {
            const unsigned u = (unsigned)(c + i) % (unsigned)(8 * p.NB);
            p.hp->win[c][i] = ((u & 7u) << 4) | ((u >> 3) * (unsigned)(kTabRows * 128));
        }
    const int smem = table_smem(p);
    if (smem > (int)prop.sharedMemPerBlockOptin) return 0;
    const bool red = p.TB != (p.ch == 2 ? kTabTB2 : p.K == TAB_F32 ? TabCfg<TAB_F32>::TB : TabCfg<TAB_F64>::TB);
    if (p.dm)
        e = p.dm == 4 ? cudaFuncSetAttribute(k_table_fir<TAB_F64, 4, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                      : cudaFuncSetAttribute(k_table_fir<TAB_F64, 4, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    else if (p.ch == 2)
        e = red ? cudaFuncSetAttribute(k_table_fir<TAB_F64, 22, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                : cudaFuncSetAttribute(k_table_fir<TAB_F64, 24, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    else
    e = p.K == TAB_F32 ? (red ? cudaFuncSetAttribute(k_table_fir<TAB_F32, 88, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                              : cudaFuncSetAttribute(k_table_fir<TAB_F32, 96, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
      : p.K == TAB_F64 ? (red ? cudaFuncSetAttribute(k_table_fir<TAB_F64, 44, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                              : cudaFuncSetAttribute(k_table_fir<TAB_F64, 48, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
                       : (red ? cudaFuncSetAttribute(k_table_fir<TAB_C64, 44, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                              : cudaFuncSetAttribute(k_table_fir<TAB_C64, 48, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (e != cudaSuccess) return (int32_t)e;
    p.ok = true;
    return 0;
}

static inline cudaError_t table_reserve(const TabPlan &p, TabRows &r, int64_t nout) {
    const int64_t rb = (int64_t)p.rowlen * p.ts;
    if (r.cap >= nout && r.row_bytes == rb) return cudaSuccess;
    tabrows_release(r);
    // two steps of slack: the kernel stages whole steps (32 rows) of the table
    cudaError_t e = cudaMalloc(&r.d_rows, (size_t)(nout + 2 * kTabStep) * (size_t)rb);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&r.d_astart, (size_t)(nout + 2 * kTabStep) * sizeof(int32_t));
    if (e != cudaSuccess) return e;
    r.cap = nout; r.row_bytes = rb;
    return cudaSuccess;
}

// One schedule slice: outputs [0, cnt) of the slice (y index y0 + k), of which the first `head` have windows that
// reach into the history.  Builds the tap rows, then launches the main kernel for [k_begin, cnt).  Returns k_begin
// (the caller computes the slice's outputs before it with the generic kernel), -1 when not covered, -2 on error.
static inline int64_t table_try_launch(TabPlan &p, TabRows &rw, const GenParams &G, int kind, int P1, int tap_is_f32,
                                       const void *d_pfb, const void *d_dpfb, const double *d_pnfb, double rate, int64_t y0, int64_t cnt,
                                       int64_t head, int64_t max_group_span, cudaStream_t st, const char **name, int64_t *launches,
                                       uint64_t tag = 0) {
    static const bool trace = getenv("MRB_TRACE") != nullptr;
#define MRB_TAB_SKIP(why) do { if (trace) fprintf(stderr, "[mrb] table kernel not used: %s\n", why); return -1; } while (0)
    if (!p.ok) MRB_TAB_SKIP("configuration not covered");
    const int es = p.es, A = p.A, BOXE = 128 / es;
    if (((uintptr_t)G.x & 15) || ((uintptr_t)G.y & 15) || (G.ldx % A) || (G.ldy % A)) MRB_TAB_SKIP("alignment");
    if (G.n_in >= (1ll << 31) - 4096 || y0 + cnt >= (1ll << 31) - 4096) MRB_TAB_SKIP("size");
    if (y0 % kTabStep) MRB_TAB_SKIP("slice start");
    if (max_group_span + p.T + A - 1 > p.rowlen) MRB_TAB_SKIP("window group wider than the tap rows");
    {   // a step's windows (32 outputs) plus two boxes of refill room must fit the ring
        const double span = (double)kTabStep / rate + p.rowlen + 2.0 * BOXE;
        if (!(rate > 0.0) || span > (double)(p.NB - 2) * BOXE) MRB_TAB_SKIP("rate too low for the ring");
    }
    const int64_t k_begin = (head + kTabStep - 1) / kTabStep * kTabStep;
    if (cnt - k_begin < 4 * kTabStep) MRB_TAB_SKIP("slice too short");
    {
        const void *before = rw.d_rows;
        if (table_reserve(p, rw, cnt) != cudaSuccess) return -2;
        if (rw.d_rows != before) rw.tag = 0;
    }

    if (tag == 0 || rw.tag != tag) {
        // pre-pass: rows + aligned starts for the whole slice (the head rows are not used); skipped when an earlier channel
        // block of the same call already built them on this stream
        rw.tag = tag;
        const unsigned g = (unsigned)ceil_div(cnt, 8);               // one warp per row
        if (p.K == TAB_F64)
            k_table_rows<double><<<g, 256, 0, st>>>((const double *)d_pfb, (const double *)d_dpfb, d_pnfb, P1, p.T, p.rowlen,
                                                    kind == 5, tap_is_f32, G.sn, G.sphi, G.salpha, G.H, cnt,
                                                    (double *)rw.d_rows, rw.d_astart, A, G.L, G.M, G.p0, G.d0m1);
        else
            k_table_rows<float><<<g, 256, 0, st>>>((const float *)d_pfb, (const float *)d_dpfb, d_pnfb, P1, p.T, p.rowlen,
                                                   kind == 5, tap_is_f32, G.sn, G.sphi, G.salpha, G.H, cnt,
                                                   (float *)rw.d_rows, rw.d_astart, A);
        ++*launches;
    }
    TabParams &P = *p.hp;
    P.k_begin = k_begin; P.N = cnt; P.y0 = y0;
    const int64_t span = cnt - k_begin;
    const int64_t groups = ceil_div(G.nch, (int64_t)kTabRows * p.ch);
    int64_t tiles = std::max<int64_t>(1, std::min<int64_t>(span / (8 * kTabStep), ceil_div((p.ch == 2 ? 4ll : 6ll * 4) * p.num_sms, groups)));
    P.KT = (int)(ceil_div(ceil_div(span, tiles), kTabStep) * kTabStep);
    tiles = ceil_div(span, P.KT);

    // complex64 moves through TMA as 8-byte elements (no arithmetic on the way)
    CUtensorMap tmx, tmy;
    const CUtensorMapDataType dt = es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t dims[2] = {(cuuint64_t)G.n_in, (cuuint64_t)G.nch};
    cuuint64_t strides[1] = {(cuuint64_t)G.ldx * es};
    cuuint32_t box[2] = {(cuuint32_t)BOXE, (cuuint32_t)(kTabRows * p.ch)};
    cuuint32_t ones[2] = {1, 1};
    if (p.encode(&tmx, dt, 2, const_cast<void *>(G.x), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        MRB_TAB_SKIP("x tensor map");
    cuuint64_t ydims[2] = {(cuuint64_t)(y0 + cnt), (cuuint64_t)G.nch};
    cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * es};
    if (p.encode(&tmy, dt, 2, G.y, ydims, ystrides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        MRB_TAB_SKIP("y tensor map");
#undef MRB_TAB_SKIP
    dim3 grid((unsigned)groups, (unsigned)tiles);
    const int smem = table_smem(p);
    const bool red = p.TB != (p.ch == 2 ? kTabTB2 : p.K == TAB_F32 ? TabCfg<TAB_F32>::TB : TabCfg<TAB_F64>::TB);
    if (p.dm == 4) {
        k_table_fir<TAB_F64, 4, 2, 4><<<grid, 128, smem, st>>>(tmx, tmy, (const double *)rw.d_rows, rw.d_astart, P);
    } else if (p.dm == 8) {
        k_table_fir<TAB_F64, 4, 2, 8><<<grid, 288, smem, st>>>(tmx, tmy, (const double *)rw.d_rows, rw.d_astart, P);
    } else if (p.ch == 2) {
        if (red) k_table_fir<TAB_F64, 22, 2><<<grid, 128 * kTabWPG, smem, st>>>(tmx, tmy, (const double *)rw.d_rows, rw.d_astart, P);
        else k_table_fir<TAB_F64, 24, 2><<<grid, 128 * kTabWPG, smem, st>>>(tmx, tmy, (const double *)rw.d_rows, rw.d_astart, P);
    } else if (p.K == TAB_F32) {
        if (red) k_table_fir<TAB_F32, 88, 1><<<grid, 128, smem, st>>>(tmx, tmy, (const float *)rw.d_rows, rw.d_astart, P);
        else k_table_fir<TAB_F32, 96, 1><<<grid, 128, smem, st>>>(tmx, tmy, (const float *)rw.d_rows, rw.d_astart, P);
    } else if (p.K == TAB_F64) {
        if (red) k_table_fir<TAB_F64, 44, 1><<<grid, 128, smem, st>>>(tmx, tmy, (const double *)rw.d_rows, rw.d_astart, P);
        else k_table_fir<TAB_F64, 48, 1><<<grid, 128, smem, st>>>(tmx, tmy, (const double *)rw.d_rows, rw.d_astart, P);
    } else {
        if (red) k_table_fir<TAB_C64, 44, 1><<<grid, 128, smem, st>>>(tmx, tmy, (const float *)rw.d_rows, rw.d_astart, P);
        else k_table_fir<TAB_C64, 48, 1><<<grid, 128, smem, st>>>(tmx, tmy, (const float *)rw.d_rows, rw.d_astart, P);
    }
    if (cudaPeekAtLastError() != cudaSuccess) return -2;
    *name = p.K == TAB_F32 ? "table_f32" : p.K == TAB_F64 ? (p.dm ? (kind <= 3 ? "int_f64_dmma" : "table_f64_dmma") : p.ch == 2 ? "table_f64_2ch" : "table_f64") : "table_c64";
    ++*launches;
    return k_begin;
}

}  // namespace mrb
