// mrb_unit.cuh -- fast path for the single-rate and interpolating kernels on real float32 samples
// (FIRStandard: src/Filters.jl:450-473, FIRInterpolator: src/Filters.jl:489-517; BASELINE configs[2]).
//
// Both read their input with unit stride: every input sample n yields L outputs y[nL + phi] = pfb[:, phi] . w(n)
// (L = 1 for FIRStandard), w(n) = the T samples ending at x[n].  All outputs of one phase share the same taps, so
// the taps cost one uniform load per R FMAs and the path is bound by the FP32 pipe (128 taps, L = 1) or by HBM
// (4 x 32 taps, L = 4), not by tap delivery:
//  * lane = channel; a CTA is 32 channels x 4 warps, warp w computes inputs [wR, wR+R) of every step for all
//    L phases: L*R accumulators per thread, FFMA with a uniform-register (or broadcast) tap operand;
//  * taps are processed in blocks of 32; per block a thread loads a window of R+32 samples from shared memory
//    (LDS.128, conflict free under SWIZZLE_128B) and reuses every sample for up to 32*L FMAs;
//  * samples arrive by TMA in [32 ch][32 samples] boxes (128-byte rows, 256-byte L2 promotion) in a ring;
//    every index is a compile-time constant relative to the step, because a step advances by exactly 4R samples;
//  * each warp stages its L*R outputs per channel in its own swizzled buffer and stores it with TMA.
// The first outputs of a chunk (window reaches into the history) are computed by k_generic.
#pragma once
#include <cstdio>

#include "mrb_tiled.cuh"

namespace mrb {

constexpr int kUnitRows = 32;           // channels per CTA
constexpr int kUnitBox = 32;            // samples per TMA box row (128 B of float32)
constexpr int kUnitBoxBytes = kUnitRows * kUnitBox * 4;   // 4096
constexpr int kUnitTB = 32;             // taps per block
constexpr int kUnitMaxBlocks = 4;       // <= 128 taps per phase
constexpr int kUnitWarps = 4;

struct alignas(16) UnitParams {
    long long n_begin, n_in;   // this launch covers inputs [n_begin, n_in)
    int KT;                    // inputs per tile (multiple of 4R)
    int nblk;                  // tap blocks per phase
    int pad0, pad1;
    // bank[phi][nblk*32]: row phi = reference pfb[:, phi] (time reversed branch), left-padded with zeros
    float bank[4 * kUnitMaxBlocks * kUnitTB];
};

template <int L, int R>
struct UnitCfg {
    static constexpr int IS = kUnitWarps * R;                       // inputs per CTA step
    static constexpr int BPS = IS / kUnitBox;                       // boxes consumed per step
    static constexpr int NB = BPS + kUnitMaxBlocks + 3;             // ring boxes: live window + 3 steps of prefetch
    static constexpr int OUT_ROW = L * R * 4;                       // bytes per channel per warp per step (64 or 128)
    static constexpr int OUT_BYTES = kUnitRows * OUT_ROW;
    static constexpr int SMEM = NB * kUnitBoxBytes + kUnitWarps * OUT_BYTES + 8 * NB;
    static_assert(IS % kUnitBox == 0 && (OUT_ROW == 64 || OUT_ROW == 128), "unsupported (L, R)");
};

template <int L, int R>
__global__ void __launch_bounds__(128, 4)
k_unit_f32(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
           const __grid_constant__ UnitParams P) {
    using C = UnitCfg<L, R>;
    constexpr int NB = C::NB;
    constexpr int NQ = (R + kUnitTB) / 4;                            // LDS.128 per window
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *out_ring = smem + NB * kUnitBoxBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(out_ring + kUnitWarps * C::OUT_BYTES);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction
    const int ch0 = blockIdx.y * kUnitRows;
    const uint32_t in_base = smem_u32(smem), bar_base = smem_u32(bars);
    const uint32_t obuf = smem_u32(out_ring) + (uint32_t)(warp * C::OUT_BYTES);
    // SWIZZLE_128B: 16-byte chunk index ^= row & 7
    const uint32_t rowpart = ((uint32_t)lane * 128u) ^ (((uint32_t)lane & 7u) << 4);

    const long long n0 = P.n_begin + (long long)blockIdx.x * P.KT;   // first input of the tile
    const int ntile = (int)min((long long)P.KT, P.n_in - n0);
    const int nsteps = (ntile + C::IS - 1) / C::IS;
    const int Tp = P.nblk * kUnitTB;
    const int xc0 = (int)(n0 - Tp);                                  // sample coordinate of box 0 (>= 0)
    const int jlast = ((nsteps * C::IS + Tp) >> 5);                  // newest box the tile touches

    if (tid == 0) {
        if (in_base & 1023u) __trap();
        for (int i = 0; i < NB; ++i) mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
        for (int jj = 0; jj < NB && jj <= jlast; ++jj) {
            mbar_expect_tx(bar_base + 8 * jj, kUnitBoxBytes);
            tma_load_2d(in_base + (uint32_t)(jj * kUnitBoxBytes), &tmx, xc0 + jj * kUnitBox, ch0, bar_base + 8 * jj);
        }
    }
    __syncthreads();

    int j_issued = min(NB, jlast + 1), i_slot = j_issued % NB;
    int j_waited = 0, w_slot = 0;
    uint32_t w_par = 0;

    for (int s = 0; s < nsteps; ++s) {
        // Two consecutive inputs of a phase share their taps, so they are accumulated as one packed pair (FFMA2 with the
        // tap as a uniform scalar operand): half the issue slots of FFMA -- the path is issue bound.  Inputs (r, r+1)
        // and tap j read window samples (1+r+j, 2+r+j), which is an aligned register pair only when r + j is odd.  Odd
        // taps therefore accumulate into pairs (0,1), (2,3), ... (accA) and even taps into pairs (-1,0), (1,2), ...,
        // (R-1,R) (accB, its two outer halves unused); the two sets meet once per step.
        unsigned long long accA[L][R / 2], accB[L][R / 2 + 1];
#pragma unroll
        for (int ph = 0; ph < L; ++ph) {
#pragma unroll
            for (int r = 0; r < R / 2; ++r) accA[ph][r] = 0ull;
#pragma unroll
            for (int r = 0; r <= R / 2; ++r) accB[ph][r] = 0ull;
        }

        for (int bb = 0; bb < P.nblk; ++bb) {
            // window of this warp for tap block bb: tile-relative samples [u0, u0 + R + 32)
            const int u0 = s * C::IS + warp * R + bb * kUnitTB;
            const int need = (u0 + R + kUnitTB - 1) >> 5;
#pragma unroll 1
            for (; j_waited <= need; ++j_waited) {
                mbar_wait(bar_base + 8 * w_slot, w_par);
                if (++w_slot == NB) { w_slot = 0; w_par ^= 1u; }
            }
            unsigned long long w2[2 * NQ];                           // w2[i] = window samples (2i, 2i+1)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int u = u0 + 4 * q;
                const uint32_t word = (uint32_t)(((u >> 2) & 7) << 4) + (uint32_t)(((u >> 5) % NB) * kUnitBoxBytes);
                const uint32_t a = in_base + (rowpart ^ word);
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w2[2 * q]), "=l"(w2[2 * q + 1]) : "r"(a) : "memory");
            }
#pragma unroll
            for (int ph = 0; ph < L; ++ph) {
                const float *taps = P.bank + (ph * P.nblk + bb) * kUnitTB;
#pragma unroll
                for (int j = 0; j < kUnitTB; j += 2) {
                    const float te = taps[j], to = taps[j + 1];
#pragma unroll
                    for (int b = 0; b <= R / 2; ++b) cfma(accB[ph][b], te, w2[b + j / 2]);        // inputs (2b-1, 2b)
#pragma unroll
                    for (int b = 0; b < R / 2; ++b) cfma(accA[ph][b], to, w2[b + j / 2 + 1]);     // inputs (2b, 2b+1)
                }
            }
        }
        float acc[L][R];
#pragma unroll
        for (int ph = 0; ph < L; ++ph) {
            float blo[R / 2 + 1], bhi[R / 2 + 1];
#pragma unroll
            for (int b = 0; b <= R / 2; ++b) asm("mov.b64 {%0, %1}, %2;" : "=f"(blo[b]), "=f"(bhi[b]) : "l"(accB[ph][b]));
#pragma unroll
            for (int b = 0; b < R / 2; ++b) {
                float alo, ahi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(alo), "=f"(ahi) : "l"(accA[ph][b]));
                acc[ph][2 * b] = alo + bhi[b];
                acc[ph][2 * b + 1] = ahi + blo[b + 1];
            }
        }

        // ---- stage this warp's L*R outputs per channel (k = n*L + phi: r-major, phase-minor) and store them
        if (lane == 0) tma_wait_read<0>();                // the previous step's store has left the buffer
        __syncwarp();
        if constexpr (L == 4) {                           // 128-byte rows, SWIZZLE_128B: one 16-byte chunk per input
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t a = obuf + (rowpart ^ (uint32_t)(r << 4));
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(acc[0][r]), "f"(acc[1][r]),
                             "f"(acc[2][r]), "f"(acc[3][r]) : "memory");
            }
        } else {                                          // 64-byte rows, SWIZZLE_64B: chunk index ^= (row >> 1) & 3
            const uint32_t rp64 = ((uint32_t)lane * 64u) ^ ((((uint32_t)lane >> 1) & 3u) << 4);
            constexpr int RPC = 4 / L;                    // inputs per 16-byte chunk
#pragma unroll
            for (int c = 0; c < R / RPC; ++c) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = acc[e % L][c * RPC + e / L];
                const uint32_t a = obuf + (rp64 ^ (uint32_t)(c << 4));
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            const long long k = (n0 + (long long)s * C::IS + warp * R) * L;   // first output of the warp's block
            tma_store_2d(&tmy, (int)k, ch0, obuf);
            tma_commit();
        }

        // ---- every warp is done with the boxes before the next step's first window: refill them
        __syncthreads();
        const int jtarget = min(((s + 1) * C::IS >> 5) + NB - 1, jlast);
        if (tid == 0) {
            int sl = i_slot;
            for (int jj = j_issued; jj <= jtarget; ++jj) {
                const uint32_t bar = bar_base + 8 * sl;
                mbar_expect_tx(bar, kUnitBoxBytes);
                tma_load_2d(in_base + (uint32_t)(sl * kUnitBoxBytes), &tmx, xc0 + jj * kUnitBox, ch0, bar);
                if (++sl == NB) sl = 0;
            }
        }
        if (jtarget >= j_issued) {
            i_slot = (i_slot + (jtarget + 1 - j_issued)) % NB;
            j_issued = jtarget + 1;
        }
    }

    for (; j_waited < j_issued; ++j_waited) {             // every issued load must have landed before exit
        mbar_wait(bar_base + 8 * w_slot, w_par);
        if (++w_slot == NB) { w_slot = 0; w_par ^= 1u; }
    }
    if (lane == 0) tma_wait_read<0>();
}

// ---------------------------------------------------------------------------------------------------------
// complex64 samples x float32 taps: the same walk with packed complex registers.  One complex x real FMA is one
// FFMA2 with the tap as a uniform-register scalar; boxes are [32 ch][16 samples] (128 bytes), taps in blocks of 16.
// (L, R) in {(1, 8), (2, 8), (4, 4)}: L*R*8 bytes of output per channel per warp per step (64 or 128).
// ---------------------------------------------------------------------------------------------------------
constexpr int kUnitCBox = 16;           // complex samples per TMA box row (128 B)
constexpr int kUnitCTB = 16;            // taps per block
constexpr int kUnitCMaxBlocks = 8;      // <= 128 taps per phase

struct alignas(16) UnitCParams {
    long long n_begin, n_in;
    int KT, nblk, pad0, pad1;
    float bank[4 * kUnitCMaxBlocks * kUnitCTB];   // bank[phi][nblk*16], left-padded with zeros
};

template <int L, int R>
struct UnitCCfg {
    static constexpr int IS = kUnitWarps * R;                       // inputs per CTA step
    // live + 3 ahead (R = 16 with one box ahead fits 3 CTAs per SM instead of 2 and measures 3 % slower)
    static constexpr int NB = (IS + kUnitCMaxBlocks * kUnitCTB + kUnitCBox - 1) / kUnitCBox + 1 + 3;
    static constexpr int OUT_ROW = L * R * 8;
    static constexpr int OUT_BYTES = kUnitRows * OUT_ROW;
    static constexpr int SMEM = NB * kUnitBoxBytes + kUnitWarps * OUT_BYTES + 8 * NB;
    static_assert(IS % kUnitCBox == 0 && (OUT_ROW == 64 || OUT_ROW == 128), "unsupported (L, R)");
};

template <int L, int R>
__global__ void __launch_bounds__(128, 3)
k_unit_c64(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
           const __grid_constant__ UnitCParams P) {
    using C = UnitCCfg<L, R>;
    constexpr int NB = C::NB, TB = kUnitCTB;
    constexpr int NQ = (R + TB) / 2;                                 // LDS.128 per window (2 samples each)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *out_ring = smem + NB * kUnitBoxBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(out_ring + kUnitWarps * C::OUT_BYTES);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ch0 = blockIdx.y * kUnitRows;
    const uint32_t in_base = smem_u32(smem), bar_base = smem_u32(bars);
    const uint32_t obuf = smem_u32(out_ring) + (uint32_t)(warp * C::OUT_BYTES);
    const uint32_t rowpart = ((uint32_t)lane * 128u) ^ (((uint32_t)lane & 7u) << 4);   // SWIZZLE_128B

    const long long n0 = P.n_begin + (long long)blockIdx.x * P.KT;
    const int ntile = (int)min((long long)P.KT, P.n_in - n0);
    const int nsteps = (ntile + C::IS - 1) / C::IS;
    const int Tp = P.nblk * TB;
    const int xc0 = (int)(n0 - Tp);                                  // sample coordinate of box 0 (>= 0, even)
    const int jlast = (nsteps * C::IS + Tp) / kUnitCBox;

    if (tid == 0) {
        if (in_base & 1023u) __trap();
#pragma unroll 1
        for (int i = 0; i < NB; ++i) mbar_init(bar_base + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmy) : "memory");
#pragma unroll 1
        for (int jj = 0; jj < NB && jj <= jlast; ++jj) {
            mbar_expect_tx(bar_base + 8 * jj, kUnitBoxBytes);
            tma_load_2d(in_base + (uint32_t)(jj * kUnitBoxBytes), &tmx, (xc0 + jj * kUnitCBox) * 2, ch0, bar_base + 8 * jj);
        }
    }
    __syncthreads();

    int j_issued = min(NB, jlast + 1), i_slot = j_issued % NB;
    int j_waited = 0, w_slot = 0;
    uint32_t w_par = 0;

    for (int s = 0; s < nsteps; ++s) {
        unsigned long long acc[L][R];
#pragma unroll
        for (int ph = 0; ph < L; ++ph)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[ph][r] = 0ull;

        for (int bb = 0; bb < P.nblk; ++bb) {
            const int u0 = s * C::IS + warp * R + bb * TB;           // tile-relative window start (even)
            const int need = (u0 + R + TB - 1) / kUnitCBox;
#pragma unroll 1
            for (; j_waited <= need; ++j_waited) {
                mbar_wait(bar_base + 8 * w_slot, w_par);
                if (++w_slot == NB) { w_slot = 0; w_par ^= 1u; }
            }
            unsigned long long w[2 * NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int u = u0 + 2 * q;
                const uint32_t word = (uint32_t)(((u >> 1) & 7) << 4) + (uint32_t)(((u / kUnitCBox) % NB) * kUnitBoxBytes);
                const uint32_t a = in_base + (rowpart ^ word);
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w[2 * q]), "=l"(w[2 * q + 1]) : "r"(a) : "memory");
            }
#pragma unroll
            for (int ph = 0; ph < L; ++ph) {
                const float *taps = P.bank + (ph * P.nblk + bb) * TB;
#pragma unroll
                for (int j = 0; j < TB; ++j) {
                    const float t = taps[j];
#pragma unroll
                    for (int r = 0; r < R; ++r) cfma(acc[ph][r], t, w[1 + r + j]);
                }
            }
        }

        // ---- stage this warp's L*R outputs per channel (k = n*L + phi) two at a time, then store them
        if (lane == 0) tma_wait_read<0>();
        __syncwarp();
        {
            const uint32_t rp = C::OUT_ROW == 128 ? rowpart : (((uint32_t)lane * 64u) ^ ((((uint32_t)lane >> 1) & 3u) << 4));
#pragma unroll
            for (int c = 0; c < L * R / 2; ++c) {
                const int o0 = 2 * c, o1 = 2 * c + 1;                // output index within the row: o = r*L + ph
                const uint32_t a = obuf + (rp ^ (uint32_t)(c << 4));
                asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a), "l"(acc[o0 % L][o0 / L]), "l"(acc[o1 % L][o1 / L]) : "memory");
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            const long long k = (n0 + (long long)s * C::IS + warp * R) * L;
            tma_store_2d(&tmy, (int)(2 * k), ch0, obuf);
            tma_commit();
        }

        __syncthreads();
        const int jtarget = min(((s + 1) * C::IS) / kUnitCBox + NB - 1, jlast);
        if (tid == 0) {
            int sl = i_slot;
#pragma unroll 1
            for (int jj = j_issued; jj <= jtarget; ++jj) {
                const uint32_t bar = bar_base + 8 * sl;
                mbar_expect_tx(bar, kUnitBoxBytes);
                tma_load_2d(in_base + (uint32_t)(sl * kUnitBoxBytes), &tmx, (xc0 + jj * kUnitCBox) * 2, ch0, bar);
                if (++sl == NB) sl = 0;
            }
        }
        if (jtarget >= j_issued) {
            i_slot = (i_slot + (jtarget + 1 - j_issued)) % NB;
            j_issued = jtarget + 1;
        }
    }
#pragma unroll 1
    for (; j_waited < j_issued; ++j_waited) {
        mbar_wait(bar_base + 8 * w_slot, w_par);
        if (++w_slot == NB) { w_slot = 0; w_par ^= 1u; }
    }
    if (lane == 0) tma_wait_read<0>();
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct UnitPlan {
    bool ok = false;
    bool cplx = false;                 // complex64 samples (k_unit_c64) instead of float32 (k_unit_f32)
    int L = 1, R = 16, nblk = 1;
    UnitParams *hp = nullptr;
    UnitCParams *hpc = nullptr;
    PFN_encodeTiled encode = nullptr;
    int num_sms = 148;
};

static inline void unit_release(UnitPlan &p) {
    delete p.hp;
    delete p.hpc;
    p.hp = nullptr;
    p.hpc = nullptr;
    p.ok = false;
}

template <int L, int R>
static inline cudaError_t unitc_set_attr() {
    return cudaFuncSetAttribute(k_unit_c64<L, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, UnitCCfg<L, R>::SMEM);
}

template <int L, int R>
static inline cudaError_t unit_set_attr() {
    return cudaFuncSetAttribute(k_unit_f32<L, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, UnitCfg<L, R>::SMEM);
}

// kind/tx/ty are the mrb.h enums (0 standard, 1 interpolator ; 0 = float32)
static inline int32_t unit_prepare(UnitPlan &p, int kind, int tx, int ty, int64_t L, int64_t M, int64_t T,
                                   const std::vector<double> &bank, const cudaDeviceProp &prop) {
    p.ok = false;
    if (!(kind == 0 || kind == 1) || tx != ty || !(tx == 0 || tx == 2) || M != 1) return 0;
    if (!(L == 1 || L == 2 || L == 4) || T > kUnitMaxBlocks * kUnitTB) return 0;
    p.cplx = tx == 2;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int32_t)(e ? e : cudaErrorUnknown);
    p.encode = (PFN_encodeTiled)fn;
    p.num_sms = prop.multiProcessorCount;
    if (p.cplx) {
        p.L = (int)L; p.R = L == 4 ? 4 : L == 1 ? 16 : 8;
        p.nblk = (int)ceil_div(T, kUnitCTB);
        p.hpc = new UnitCParams();
        memset(p.hpc, 0, sizeof(UnitCParams));
        p.hpc->nblk = p.nblk;
        const int64_t Tpc = (int64_t)p.nblk * kUnitCTB;
        for (int64_t ph = 0; ph < L; ++ph)
            for (int64_t i = 0; i < T; ++i) p.hpc->bank[ph * Tpc + (Tpc - T) + i] = (float)bank[ph * T + i];
        e = L == 1 ? unitc_set_attr<1, 16>() : L == 2 ? unitc_set_attr<2, 8>() : unitc_set_attr<4, 4>();
        if (e != cudaSuccess) return (int32_t)e;
        p.ok = true;
        return 0;
    }
    p.L = (int)L; p.R = L == 1 ? 16 : 8;
    p.nblk = (int)ceil_div(T, kUnitTB);
    p.hp = new UnitParams();
    memset(p.hp, 0, sizeof(UnitParams));
    p.hp->nblk = p.nblk;
    const int64_t Tp = (int64_t)p.nblk * kUnitTB;
    for (int64_t ph = 0; ph < L; ++ph)
        for (int64_t i = 0; i < T; ++i) p.hp->bank[ph * Tp + (Tp - T) + i] = (float)bank[ph * T + i];
    e = L == 1 ? unit_set_attr<1, 16>() : L == 2 ? unit_set_attr<2, 8>() : unit_set_attr<4, 8>();
    if (e != cudaSuccess) return (int32_t)e;
    p.ok = true;
    return 0;
}

// Live tap update: rewrite the padded bank of an existing plan (host parameter block; the next launch carries it).
static inline void unit_set_bank(UnitPlan &p, int64_t T, const std::vector<double> &bank) {
    if (!p.ok) return;
    if (p.cplx) {
        const int64_t Tpc = (int64_t)p.nblk * kUnitCTB;
        for (int64_t ph = 0; ph < p.L; ++ph)
            for (int64_t i = 0; i < T; ++i) p.hpc->bank[ph * Tpc + (Tpc - T) + i] = (float)bank[ph * T + i];
    } else {
        const int64_t Tp = (int64_t)p.nblk * kUnitTB;
        for (int64_t ph = 0; ph < p.L; ++ph)
            for (int64_t i = 0; i < T; ++i) p.hp->bank[ph * Tp + (Tp - T) + i] = (float)bank[ph * T + i];
    }
}

// Launch for inputs [n_begin, n_in) of this chunk.  Returns the first OUTPUT the kernel covers (>= 0; the caller
// computes the outputs before it with the generic kernel), -1 when the call is not covered, -2 on a CUDA error.
static inline int64_t unit_try_launch(UnitPlan &p, const GenParams &G, cudaStream_t st, const char **name,
                                      int64_t *launches) {
    static const bool trace = getenv("MRB_TRACE") != nullptr;
#define MRB_UNIT_SKIP(why) do { if (trace) fprintf(stderr, "[mrb] unit kernel not used: %s\n", why); return -1; } while (0)
    if (!p.ok) MRB_UNIT_SKIP("configuration not covered");
    if (G.mode != SEQ_INTEGER || G.p0 != 0 || G.d0m1 != 0) MRB_UNIT_SKIP("carried phase/deficit");
    const int al = p.cplx ? 1 : 3;
    if (((uintptr_t)G.x & 15) || ((uintptr_t)G.y & 15) || (G.ldx & al) || (G.ldy & al)) MRB_UNIT_SKIP("alignment");
    if (G.n_in >= (1ll << 30) - 4096 || G.nout >= (1ll << 30) - 4096) MRB_UNIT_SKIP("size");
    if (p.cplx) {
        const int64_t Tp = (int64_t)p.nblk * kUnitCTB;
        const int64_t n_begin = (Tp + kUnitCBox - 1) / kUnitCBox * kUnitCBox;   // padded window inside x, box aligned
        const int IS = kUnitWarps * p.R;
        if (G.n_in - n_begin < 4 * IS) MRB_UNIT_SKIP("chunk too short");
        UnitCParams &P = *p.hpc;
        P.n_begin = n_begin; P.n_in = G.n_in;
        const int64_t span = G.n_in - n_begin;
        const int64_t groups = ceil_div(G.nch, kUnitRows);
        int64_t tiles = std::max<int64_t>(1, std::min<int64_t>(span / (8 * IS), ceil_div(12ll * 3 * p.num_sms, groups)));
        P.KT = (int)(ceil_div(ceil_div(span, tiles), IS) * IS);
        tiles = ceil_div(span, P.KT);
        CUtensorMap tmx, tmy;
        cuuint64_t dims[2] = {(cuuint64_t)(2 * G.n_in), (cuuint64_t)G.nch};
        cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 8};
        cuuint32_t box[2] = {2 * kUnitCBox, kUnitRows};
        cuuint32_t es[2] = {1, 1};
        if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            MRB_UNIT_SKIP("x tensor map");
        cuuint64_t ydims[2] = {(cuuint64_t)(2 * G.nout), (cuuint64_t)G.nch};
        cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * 8};
        cuuint32_t ybox[2] = {(cuuint32_t)(2 * p.L * p.R), kUnitRows};
        if (p.encode(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G.y, ydims, ystrides, ybox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     p.L * p.R * 8 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            MRB_UNIT_SKIP("y tensor map");
        dim3 grid((unsigned)tiles, (unsigned)groups);
        if (p.L == 1) k_unit_c64<1, 16><<<grid, 128, UnitCCfg<1, 16>::SMEM, st>>>(tmx, tmy, P);
        else if (p.L == 2) k_unit_c64<2, 8><<<grid, 128, UnitCCfg<2, 8>::SMEM, st>>>(tmx, tmy, P);
        else k_unit_c64<4, 4><<<grid, 128, UnitCCfg<4, 4>::SMEM, st>>>(tmx, tmy, P);
        if (cudaPeekAtLastError() != cudaSuccess) return -2;
        *name = p.L == 1 ? "unit_c64_l1_r16" : p.L == 2 ? "unit_c64_l2_r8" : "unit_c64_l4_r4";
        ++*launches;
        return n_begin * p.L;
    }
    const int64_t Tp = (int64_t)p.nblk * kUnitTB;
    const int64_t n_begin = Tp;                                // first input whose padded window lies inside x
    const int IS = kUnitWarps * p.R;
    if (G.n_in - n_begin < 4 * IS) MRB_UNIT_SKIP("chunk too short");

    UnitParams &P = *p.hp;
    P.n_begin = n_begin; P.n_in = G.n_in;
    const int64_t span = G.n_in - n_begin;
    // time tiles: whole steps, enough CTAs to fill the machine a few times over
    const int64_t groups = ceil_div(G.nch, kUnitRows);
    static const int wv = getenv("MRB_UNIT_WAVES") ? atoi(getenv("MRB_UNIT_WAVES")) : 12;
    int64_t tiles = std::max<int64_t>(1, std::min<int64_t>(span / (8 * IS), ceil_div((int64_t)wv * 4 * p.num_sms, groups)));
    P.KT = (int)(ceil_div(ceil_div(span, tiles), IS) * IS);
    tiles = ceil_div(span, P.KT);

    CUtensorMap tmx, tmy;
    cuuint64_t dims[2] = {(cuuint64_t)G.n_in, (cuuint64_t)G.nch};
    cuuint64_t strides[1] = {(cuuint64_t)G.ldx * 4};
    cuuint32_t box[2] = {kUnitBox, kUnitRows};
    cuuint32_t es[2] = {1, 1};
    if (p.encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(G.x), dims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        MRB_UNIT_SKIP("x tensor map");
    cuuint64_t ydims[2] = {(cuuint64_t)G.nout, (cuuint64_t)G.nch};
    cuuint64_t ystrides[1] = {(cuuint64_t)G.ldy * 4};
    cuuint32_t ybox[2] = {(cuuint32_t)(p.L * p.R), kUnitRows};
    if (p.encode(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G.y, ydims, ystrides, ybox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 p.L == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        MRB_UNIT_SKIP("y tensor map");
#undef MRB_UNIT_SKIP
    dim3 grid((unsigned)tiles, (unsigned)groups);
    if (p.L == 1) k_unit_f32<1, 16><<<grid, 128, UnitCfg<1, 16>::SMEM, st>>>(tmx, tmy, P);
    else if (p.L == 2) k_unit_f32<2, 8><<<grid, 128, UnitCfg<2, 8>::SMEM, st>>>(tmx, tmy, P);
    else k_unit_f32<4, 8><<<grid, 128, UnitCfg<4, 8>::SMEM, st>>>(tmx, tmy, P);
    if (cudaPeekAtLastError() != cudaSuccess) return -2;
    *name = p.L == 1 ? "unit_f32_l1_r16" : p.L == 2 ? "unit_f32_l2_r8" : "unit_f32_l4_r8";
    ++*launches;
    return n_begin * p.L;
}

}  // namespace mrb
