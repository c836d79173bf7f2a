// mma_probe.cu -- stand-alone probe of the tcgen05 mechanics the tensor-core FIR kernel (csrc/mrb_mma.cuh) relies on.
// Not part of the library.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe mma_probe.cu
// Run on a B200: prints one line per check.
//   T1  tcgen05.st -> tcgen05.ld round trip (lane / column addressing of the 32x32b shape)
//   T2  kind::tf32 MMA, A from TENSOR MEMORY (lane = row, one tf32 per column), B from shared memory (K-major,
//       SWIZZLE_128B canonical layout), D in tensor memory: exact on tf32-representable integers
//   T3  the same with A from shared memory (SS mode)
//   T4  3xTF32 split (hi*hi + hi*lo + lo*hi) on random float32 data against a float64 reference
//   T5  issue rate: cycles per MMA for N = 16..128, TS and SS mode
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(lbo16 & 0x3fff) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;                     // version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

constexpr int KMAX = 128;            // K extent of the test matrices (4 swizzle atoms of 32 floats)
constexpr int NMAX = 128;

// shared-memory image offset (bytes) of element (row, k) of a K-major SWIZZLE_128B matrix with `rows` rows
__host__ __device__ inline uint32_t sw128_off(int rows, int row, int k) {
    const int atom = k >> 5, kk = k & 31, chunk = kk >> 2, e = kk & 3;
    return (uint32_t)(atom * rows * 128 + row * 128 + ((chunk ^ (row & 7)) << 4) + e * 4);
}

struct ProbeArgs {
    const float *A_hi, *A_lo;        // [128][KMAX] row-major
    const float *Bimg_hi, *Bimg_lo;  // shared-memory images [KMAX/32][N][32] swizzled
    const float *Aimg_hi;            // SS mode: image [KMAX/32][128][32] swizzled
    float *D;                        // [128][N]
    float *RT;                       // round trip [128][KMAX]
    long long *cycles;               // timing results
    int N, mode, lbo, reps;          // mode 0: round trip; 1: TS single pass; 2: SS single pass; 3: TS 3xTF32; 4: timing TS; 5: timing SS
};

__global__ void __launch_bounds__(160, 1) k_probe(const ProbeArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *b_hi = smem;                       // 4 atoms x 128 rows x 128 B = 64 KB (N <= 128)
    unsigned char *b_lo = smem + 65536;
    unsigned char *a_hi = smem + 131072;              // 64 KB
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = P.N;

    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // images -> shared memory (generic proxy), then make them visible to the async proxy (tensor core)
    {
        const int nb = KMAX / 32 * N * 128 / 16;
        for (int i = tid; i < nb; i += blockDim.x) {
            reinterpret_cast<uint4 *>(b_hi)[i] = reinterpret_cast<const uint4 *>(P.Bimg_hi)[i];
            reinterpret_cast<uint4 *>(b_lo)[i] = reinterpret_cast<const uint4 *>(P.Bimg_lo)[i];
        }
        const int na = KMAX / 32 * 128 * 128 / 16;
        for (int i = tid; i < na; i += blockDim.x) reinterpret_cast<uint4 *>(a_hi)[i] = reinterpret_cast<const uint4 *>(P.Aimg_hi)[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t colD = 0, colAh = 128, colAl = 256;          // D: columns 0..127, A_hi 128..255, A_lo 256..383

    if (warp < 4) {
        // thread = row (TMEM lane 32*warp + lane); its K values go to consecutive columns
        const int row = warp * 32 + lane;
        const uint32_t lanebase = tb + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < KMAX; c += 8) {
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                vh[e] = __float_as_uint(P.A_hi[row * KMAX + c + e]);
                vl[e] = __float_as_uint(P.A_lo[row * KMAX + c + e]);
            }
            st32(lanebase + colAh + c, vh);
            st32(lanebase + colAl + c, vl);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (P.mode == 0) {
            for (int c = 0; c < KMAX; c += 8) {
                uint32_t v[8];
                ld32(lanebase + colAh + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 8; ++e) P.RT[row * KMAX + c + e] = __uint_as_float(v[e]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (P.mode >= 1) {
        const uint32_t idesc = make_idesc(128, N);
        if (tid == 128) {                                       // warp 4, lane 0: the MMA issuer
            long long t0 = clock64();
            const int reps = P.mode >= 4 ? P.reps : 1;
            for (int r = 0; r < reps; ++r) {
                for (int ks = 0; ks < KMAX / 8; ++ks) {
                    const uint32_t boff = (uint32_t)((ks >> 2) * N * 128 + (ks & 3) * 32);
                    const uint64_t bh = make_desc(smem_u32(b_hi) + boff, (uint32_t)P.lbo);
                    const uint64_t bl = make_desc(smem_u32(b_lo) + boff, (uint32_t)P.lbo);
                    const uint32_t acc = (ks > 0 || r > 0) ? 1u : 0u;
                    if (P.mode == 1 || P.mode == 3 || P.mode == 4) {
                        mma_ts(tb + colD, tb + colAh + 8 * ks, bh, idesc, acc);
                        if (P.mode != 1) {
                            mma_ts(tb + colD, tb + colAh + 8 * ks, bl, idesc, 1u);
                            mma_ts(tb + colD, tb + colAl + 8 * ks, bh, idesc, 1u);
                        }
                    } else {
                        const uint32_t aoff = (uint32_t)((ks >> 2) * 128 * 128 + (ks & 3) * 32);
                        const uint64_t ah = make_desc(smem_u32(a_hi) + aoff, (uint32_t)P.lbo);
                        mma_ss(tb + colD, ah, bh, idesc, acc);
                        if (P.mode == 5) {
                            mma_ss(tb + colD, ah, bl, idesc, 1u);
                            mma_ss(tb + colD, ah, bh, idesc, 1u);
                        }
                    }
                }
            }
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            long long t1 = clock64();
            if (P.cycles) P.cycles[0] = t1 - t0;
        }
        __syncthreads();
        if (warp < 4) {
            mbar_wait(smem_u32(&bar), 0);
            tc_fence_after();
            const int row = warp * 32 + lane;
            const uint32_t lanebase = tb + ((uint32_t)(warp * 32) << 16);
            for (int c = 0; c < N; c += 8) {
                uint32_t v[8];
                ld32(lanebase + colD + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 8; ++e) P.D[row * N + c + e] = __uint_as_float(v[e]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// Issue-rate probe: the whole warp enters, one ELECTED lane issues (ptxas then keeps the operands in uniform
// registers without a divergence waterfall).  NACC independent accumulators are used round robin.
// Variants (flags): 1 = tcgen05.commit after every 48 MMAs (one "group", as the FIR kernel does);
//   2 = the FIR kernel's tensor-memory layout (D at columns 0..63, A rings at 64.. and 288..);
//   4 = warps 0-3 hammer tcgen05.st into unrelated columns meanwhile;  8 = warps 0-3 hammer tcgen05.ld of D meanwhile;
//   16 = non-zero data in B (shared memory) and A (tensor memory).
template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(160, 1) k_rate(long long *cycles, int reps, int flags) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) unsigned long long bar2;
    __shared__ volatile int stop;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    for (int i = tid; i < 3 * 65536 / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem)[i] = (flags & 16) ? make_uint4(0x3f800000u + 8192u * (i & 63), 0x3f000000u, 0xbf800000u + 8192u * (i & 7), 0x3e000000u)
                                                           : make_uint4(0, 0, 0, 0);
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        stop = 0;
        mbar_init(smem_u32(&bar), 1);
        mbar_init(smem_u32(&bar2), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t colD = 0, colAh = (flags & 2) ? 64 : 256, colAl = (flags & 2) ? 288 : 384;
    if (warp < 4 && (flags & 16)) {
        const uint32_t lanebase = tb + ((uint32_t)(warp * 32) << 16);
        for (int c = 64; c < 512; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0x3f800000u + 8192u * (uint32_t)((c + e + lane) & 63);
            st32(lanebase + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4) {
        const uint32_t idesc = make_idesc(128, N);
        const uint32_t sb = smem_u32(smem);
        const uint64_t bdesc0 = make_desc(sb, 1), adesc0 = make_desc(sb + 131072, 1);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const uint32_t boff = (uint32_t)((ks >> 2) * N * 128 + (ks & 3) * 32) >> 4;
                const uint32_t aoff = (uint32_t)((ks >> 2) * 128 * 128 + (ks & 3) * 32) >> 4;
                const uint32_t d = tb + colD + (uint32_t)((ks % NACC) * N);
                if (elect_one()) {
                    if (TS) {
                        mma_ts(d, tb + colAh + 8 * ks, bdesc0 + boff, idesc, 1u);
                        mma_ts(d, tb + colAh + 8 * ks, bdesc0 + boff + 4096, idesc, 1u);
                        mma_ts(d, tb + colAl + 8 * ks, bdesc0 + boff, idesc, 1u);
                    } else {
                        mma_ss(d, adesc0 + aoff, bdesc0 + boff, idesc, 1u);
                        mma_ss(d, adesc0 + aoff, bdesc0 + boff + 4096, idesc, 1u);
                        mma_ss(d, adesc0 + aoff, bdesc0 + boff, idesc, 1u);
                    }
                }
                __syncwarp();
            }
            if ((flags & 1) && r + 1 < reps) {
                if (elect_one()) tc_commit(smem_u32(&bar2));          // nobody waits on bar2: its phase just keeps flipping
                __syncwarp();
            }
        }
        if (elect_one()) tc_commit(smem_u32(&bar));
        __syncwarp();
        // with per-group commits the barrier has flipped many times: wait for the LAST MMA by polling both parities
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (tid == 128) cycles[0] = t1 - t0;
        if (tid == 128) stop = 1;
    } else if (flags & (32 | 64)) {
        // shared-memory traffic from the other warps: 32 = conflict-free LDS.128 + STS.128 on another region (what the
        // converter / epilogue warps of the FIR kernel do); 64 = only every 8th pass (lighter)
        unsigned char *scratch = smem + 160 * 1024;
        const uint32_t base = smem_u32(scratch) + (uint32_t)(tid * 16);
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        int it = 0;
        while (!stop) {
            if ((flags & 64) && (++it & 7)) { __nanosleep(40); continue; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 w;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(base + (uint32_t)(i * 2048)) : "memory");
                v.x += w.x; v.y += w.y;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(i * 2048 + 16384)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        }
        if (v.x == 12345.f) cycles[1] = 1;
    } else if (flags & (4 | 8)) {
        const uint32_t lanebase = tb + ((uint32_t)(warp * 32) << 16);
        uint32_t v[8] = {1, 2, 3, 4, 5, 6, 7, 8};
        while (!stop) {
            if (flags & 4) { st32(lanebase + 200, v); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
            if (flags & 8) { ld32(lanebase + 32, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
        }
        if (v[0] == 0xdeadbeef) cycles[1] = v[1];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int N, bool TS, int NACC>
static int rate(long long *dcyc, const char *what, int flags = 0) {
    const int SMEM = 3 * 65536, reps = 64;
    if (cudaFuncSetAttribute(k_rate<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return 1;
    long long cyc = 0;
    for (int it = 0; it < 2; ++it) {
        k_rate<N, TS, NACC><<<1, 160, SMEM>>>(dcyc, reps, flags);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("k_rate failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("T6 %s N=%3d accumulators=%d flags=%2d: %.1f cycles per MMA (M128 x N x K8, %d MMAs)\n", what, N, NACC, flags, (double)cyc / (reps * 48), reps * 48);
    return 0;
}

static float tf32_rna(float x) {                 // round to nearest (ties away) to 10 explicit mantissa bits
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    const int SMEM = 3 * 65536;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    std::vector<float> A(128 * KMAX), Ah(128 * KMAX), Al(128 * KMAX);
    std::vector<float> B(NMAX * KMAX), Bh(NMAX * KMAX), Bl(NMAX * KMAX);
    float *dAh, *dAl, *dBh, *dBl, *dAimg, *dD, *dRT;
    long long *dcyc;
    CK(cudaMalloc(&dAh, A.size() * 4)); CK(cudaMalloc(&dAl, A.size() * 4));
    CK(cudaMalloc(&dBh, 65536)); CK(cudaMalloc(&dBl, 65536)); CK(cudaMalloc(&dAimg, 65536));
    CK(cudaMalloc(&dD, 128 * NMAX * 4)); CK(cudaMalloc(&dRT, 128 * KMAX * 4)); CK(cudaMalloc(&dcyc, 8));

    auto run = [&](int mode, int N, int lbo, bool exact_ints, int reps, double *err, long long *cyc) -> int {
        srand(1234 + mode + N);
        for (size_t i = 0; i < A.size(); ++i) {
            A[i] = exact_ints ? (float)(rand() % 17 - 8) : (float)rand() / RAND_MAX;
            Ah[i] = exact_ints ? A[i] : tf32_rna(A[i]);
            Al[i] = exact_ints ? 0.f : tf32_rna(A[i] - Ah[i]);
        }
        for (size_t i = 0; i < B.size(); ++i) {
            B[i] = exact_ints ? (float)(rand() % 9 - 4) : (float)rand() / RAND_MAX - 0.5f;
            Bh[i] = exact_ints ? B[i] : tf32_rna(B[i]);
            Bl[i] = exact_ints ? 0.f : tf32_rna(B[i] - Bh[i]);
        }
        std::vector<float> imgh(16384, 0.f), imgl(16384, 0.f), imga(16384, 0.f);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < KMAX; ++k) {
                imgh[sw128_off(N, n, k) / 4] = Bh[n * KMAX + k];
                imgl[sw128_off(N, n, k) / 4] = Bl[n * KMAX + k];
            }
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < KMAX; ++k) imga[sw128_off(128, m, k) / 4] = Ah[m * KMAX + k];
        CK(cudaMemcpy(dAh, Ah.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAl, Al.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBh, imgh.data(), 65536, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBl, imgl.data(), 65536, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAimg, imga.data(), 65536, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0, 128 * NMAX * 4));
        ProbeArgs P{dAh, dAl, dBh, dBl, dAimg, dD, dRT, dcyc, N, mode, lbo, reps};
        k_probe<<<1, 160, SMEM>>>(P);
        CK(cudaDeviceSynchronize());
        if (cyc) CK(cudaMemcpy(cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        if (mode == 0) {
            std::vector<float> rt(128 * KMAX);
            CK(cudaMemcpy(rt.data(), dRT, rt.size() * 4, cudaMemcpyDeviceToHost));
            double e = 0;
            for (size_t i = 0; i < rt.size(); ++i) e = fmax(e, fabs((double)rt[i] - Ah[i]));
            *err = e;
            return 0;
        }
        std::vector<float> D(128 * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double e = 0, mx = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < KMAX; ++k)
                    ref += (mode == 1 || mode == 2) ? (double)Ah[m * KMAX + k] * Bh[n * KMAX + k] : (double)A[m * KMAX + k] * B[n * KMAX + k];
                e = fmax(e, fabs(ref - D[m * N + n]));
                mx = fmax(mx, fabs(ref));
            }
        *err = e / mx;
        return 0;
    };

    double err;
    long long cyc;
    if (run(0, 32, 1, true, 1, &err, nullptr)) return 1;
    printf("T1 tmem st/ld round trip: max abs err %.3g\n", err);
    for (int lbo = 0; lbo <= 1; ++lbo)
        for (int N : {16, 32, 64, 128}) {
            if (run(1, N, lbo, true, 1, &err, nullptr)) return 1;
            printf("T2 TS-mode tf32 exact ints  N=%3d lbo=%d: normalised err %.3g\n", N, lbo, err);
        }
    for (int N : {16, 32, 128}) {
        if (run(2, N, 1, true, 1, &err, nullptr)) return 1;
        printf("T3 SS-mode tf32 exact ints  N=%3d: normalised err %.3g\n", N, err);
    }
    for (int N : {16, 32, 64}) {
        if (run(1, N, 1, false, 1, &err, nullptr)) return 1;
        printf("T4a single-pass tf32 on random f32 (hi only) N=%3d: normalised err %.3g\n", N, err);
        if (run(3, N, 1, false, 1, &err, nullptr)) return 1;
        printf("T4b 3xTF32 on random f32              N=%3d: normalised err %.3g\n", N, err);
    }
    for (int N : {16, 32, 64, 128}) {
        const int reps = 64;
        if (run(4, N, 1, false, reps, &err, &cyc)) return 1;
        printf("T5 TS timing N=%3d: %lld cycles for %d MMAs (M128 K8) = %.1f cycles/MMA\n", N, cyc, reps * 16 * 3, (double)cyc / (reps * 16 * 3));
        if (run(5, N, 1, false, reps, &err, &cyc)) return 1;
        printf("T5 SS timing N=%3d: %lld cycles for %d MMAs (M128 K8) = %.1f cycles/MMA\n", N, cyc, reps * 16 * 3, (double)cyc / (reps * 16 * 3));
    }
    if (rate<32, true, 1>(dcyc, "TS", 0) || rate<32, true, 1>(dcyc, "TS", 32) || rate<32, true, 1>(dcyc, "TS", 64) ||
        rate<32, true, 1>(dcyc, "TS", 16 + 32) || rate<64, true, 1>(dcyc, "TS", 32) || rate<16, true, 1>(dcyc, "TS", 32) ||
        rate<128, true, 1>(dcyc, "TS", 32) || rate<32, false, 1>(dcyc, "SS", 32))
        return 1;
    return 0;
}
