// proto_table2ch.cu -- COMPILE-ONLY feasibility prototype for DESIGN.md section 8 item 1(a): the arbitrary / Farrow
// table kernel with TWO channels per lane.  Not part of libmrb.so, never launched by the product; it exists to answer
// "does the inner loop fit the register file at 3 CTAs x 128 threads per SM (168 registers), and what is its
// instruction mix?" before the kernel is rebuilt around it.  Build and inspect:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -Xptxas -v -o /tmp/p.cubin tools/proto_table2ch.cu
//   cuobjdump -sass /tmp/p.cubin | grep -c FFMA2
//
// Per warp and window group (8 outputs), per tap block of TB = 44 elements:
//   window: 2 channels x 11 LDS.128 (conflict free, SWIZZLE_128B rows of a [64 ch][32 samples] box)
//   taps:   8 outputs x 11 warp-uniform LDS.128 -- each now feeds 2 channels x 2 FFMA2 instead of 1 x 2
// so the broadcast tap loads per FMA halve (the current kernel is bound by them: shared-memory wavefronts at 90 %).
#include <cstdint>

constexpr int TB = 44, NQ = TB / 4, OPW = 8, NBLK = 2;

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

// smem layout of the prototype: [0, 64 KB) ring of 8 boxes [64 ch][128 B], then the step's 32 tap rows [32][88] floats
__global__ void __launch_bounds__(128, 3)
k_proto_table2ch(const float *__restrict__ x, const float *__restrict__ rows, float *__restrict__ y, const int *__restrict__ astart,
                 int nsteps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *rows_s = reinterpret_cast<float *>(smem + 8 * 8192);
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t in_base = (uint32_t)__cvta_generic_to_shared(smem);
    // SWIZZLE_128B: chunk ^= row & 7; rows lane and lane + 32 share the XOR term
    const uint32_t rp0 = ((uint32_t)lane * 128u) ^ (((uint32_t)lane & 7u) << 4);
    const uint32_t rp1 = rp0 + 32u * 128u;
    for (int s = 0; s < nsteps; ++s) {
        // (staging stands in for the TMA ring and the bulk copies of the real kernel)
        for (int i = threadIdx.x; i < 8 * 8192 / 16; i += 128)
            reinterpret_cast<float4 *>(smem)[i] = reinterpret_cast<const float4 *>(x)[(size_t)s * 4096 + i];
        for (int i = threadIdx.x; i < 32 * NBLK * TB / 4; i += 128)
            reinterpret_cast<float4 *>(rows_s)[i] = reinterpret_cast<const float4 *>(rows)[(size_t)s * 32 * NBLK * TB / 4 + i];
        __syncthreads();
        const int a0 = astart[s * 4 + warp];                            // aligned window start (samples), uniform
        unsigned long long acc[OPW][2];                                 // [output][channel]: (even taps, odd taps) sums
#pragma unroll
        for (int o = 0; o < OPW; ++o) acc[o][0] = acc[o][1] = 0ull;
        const float *rowg = rows_s + warp * OPW * NBLK * TB;
        for (int bb = 0; bb < NBLK; ++bb) {
            float w0[TB], w1[TB];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int u = (a0 + bb * TB) / 4 + q;                   // 16-byte chunk index in the ring
                const uint32_t word = (uint32_t)((u & 7) << 4) + (uint32_t)(((u >> 3) & 7) * 8192);
                const uint32_t ad0 = in_base + (rp0 ^ word), ad1 = in_base + (rp1 ^ word);
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(w0[4 * q]), "=f"(w0[4 * q + 1]), "=f"(w0[4 * q + 2]), "=f"(w0[4 * q + 3]) : "r"(ad0) : "memory");
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(w1[4 * q]), "=f"(w1[4 * q + 1]), "=f"(w1[4 * q + 2]), "=f"(w1[4 * q + 3]) : "r"(ad1) : "memory");
            }
#pragma unroll
            for (int o = 0; o < OPW; ++o) {
                const float *tr = rowg + o * NBLK * TB + bb * TB;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const float4 t = reinterpret_cast<const float4 *>(tr)[q];
                    const unsigned long long t01 = pk(t.x, t.y), t23 = pk(t.z, t.w);
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[o][0]) : "l"(t01), "l"(pk(w0[4 * q], w0[4 * q + 1])));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[o][1]) : "l"(t01), "l"(pk(w1[4 * q], w1[4 * q + 1])));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[o][0]) : "l"(t23), "l"(pk(w0[4 * q + 2], w0[4 * q + 3])));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[o][1]) : "l"(t23), "l"(pk(w1[4 * q + 2], w1[4 * q + 3])));
                }
            }
        }
#pragma unroll
        for (int o = 0; o < OPW; ++o)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[o][c]));
                y[((size_t)(s * 4 + warp) * OPW + o) * 64 + c * 32 + lane] = lo + hi;
            }
        __syncthreads();
    }
}
