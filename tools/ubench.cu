// ubench.cu -- pipe-rate microbenchmarks that size the tiled FIR kernels (run on the B200 box):
//   FFMA vs FFMA2 (fma.rn.f32x2) issue rate, LDS.32/.64/.128 broadcast and strided rates, DFMA rate.
// Prints warp-instructions per clock per SM, from clock64() deltas of a full-occupancy launch.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in, long long *cyc) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = in[i & 1023];
    __syncthreads();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = in[threadIdx.x + i];
    float b = in[5], c = in[7];
    long long t0 = clock64();
    if (MODE == 0) {          // FFMA, 16 independent chains
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        }
    } else if (MODE == 1) {   // FFMA2
        unsigned long long bb, cc;
        asm("mov.b64 %0, {%1,%2};" : "=l"(bb) : "f"(b), "f"(b));
        asm("mov.b64 %0, {%1,%2};" : "=l"(cc) : "f"(c), "f"(c));
        unsigned long long v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(v[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(bb), "l"(cc));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0,%1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(v[i]));
    } else if (MODE == 2) {   // LDS.32 broadcast (all lanes same address), 16 per iter
        int off = (int)(in[3]) & 7;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] += sm[off + i * 4 + (it & 63) * 64];
        }
    } else if (MODE == 3) {   // LDS.128 broadcast
        int off = ((int)(in[3]) & 7) * 4;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = *reinterpret_cast<const float4 *>(&sm[off + i * 4 + (it & 63) * 64]);
                a[4 * i] += v.x; a[4 * i + 1] += v.y; a[4 * i + 2] += v.z; a[4 * i + 3] += v.w;
            }
        }
    } else if (MODE == 4) {   // LDS.64 lane-strided, conflict free (lane*2 words), 8 per iter
        int lane = threadIdx.x & 31;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = *reinterpret_cast<const float2 *>(&sm[lane * 2 + i * 64 + (it & 7) * 512]);
                a[2 * i] += v.x; a[2 * i + 1] += v.y;
            }
        }
    } else if (MODE == 5) {   // LDS.64 with row pitch 66 words (lane = row): the transposed read of the tiled kernel
        int lane = threadIdx.x & 31;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = *reinterpret_cast<const float2 *>(&sm[lane * 66 + i * 2 + (it & 7) * 16]);
                a[2 * i] += v.x; a[2 * i + 1] += v.y;
            }
        }
    } else if (MODE == 6) {   // LDS.128 with row pitch 68 words (lane = row)
        int lane = threadIdx.x & 31;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = *reinterpret_cast<const float4 *>(&sm[lane * 68 + i * 4 + (it & 7) * 16]);
                a[4 * i] += v.x; a[4 * i + 1] += v.y; a[4 * i + 2] += v.z; a[4 * i + 3] += v.w;
            }
        }
    } else if (MODE == 7) {   // DFMA
        double d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = a[i];
        double db = b, dc = c;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] = fma(d[i], db, dc);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (float)d[i];
    } else if (MODE == 8) {   // FFMA2 + LDS.32 broadcast interleaved 1:1 (the tiled inner loop's mix)
        unsigned long long v[8], cc;
        asm("mov.b64 %0, {%1,%2};" : "=l"(cc) : "f"(c), "f"(c));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(v[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        int off = (int)(in[3]) & 7;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float t = sm[off + i * 4 + (it & 63) * 64];
                unsigned long long tt;
                asm("mov.b64 %0, {%1,%2};" : "=l"(tt) : "f"(t), "f"(t));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(tt), "l"(cc));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float t = sm[off + i * 4 + 32 + (it & 63) * 64];
                unsigned long long tt;
                asm("mov.b64 %0, {%1,%2};" : "=l"(tt) : "f"(t), "f"(t));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(tt), "l"(cc));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0,%1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(v[i]));
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double instr_per_iter_per_warp, int blocks_per_sm) {
    int nsm = 148;
    float *out, *in; long long *cyc;
    int nb = nsm * blocks_per_sm;
    cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, nb * 8);
    cudaMemset(in, 0, 4096 * 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<nb, 256, 32768>>>(out, in, cyc);
    cudaEventRecord(e0);
    k<MODE><<<nb, 256, 32768>>>(out, in, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = new long long[nb]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; ++i) avg += h[i]; avg /= nb;
    double warp_instr_per_sm = (double)ITERS * instr_per_iter_per_warp * 8 * blocks_per_sm;
    printf("%-34s blocks/SM=%d  cycles=%.0f  warp-instr/clk/SM=%.3f  time=%.3f ms  (%.1f Ginstr/s chip)  err=%s\n", name,
           blocks_per_sm, avg, warp_instr_per_sm / avg, ms, warp_instr_per_sm * nsm / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(in); cudaFree(cyc); delete[] h;
}

int main() {
    for (int b : {1, 2, 4}) {
        run<0>("FFMA (3-reg)", 16, b);
        run<1>("FFMA2 (fma.rn.f32x2)", 16, b);
        run<2>("LDS.32 broadcast", 16, b);
        run<3>("LDS.128 broadcast", 4, b);
        run<4>("LDS.64 lane-contiguous", 8, b);
        run<5>("LDS.64 row pitch 66w (transposed)", 8, b);
        run<6>("LDS.128 row pitch 68w (transposed)", 4, b);
        run<7>("DFMA", 16, b);
        run<8>("FFMA2 + LDS.32 bcast 1:1 (count FFMA2)", 16, b);
    }
    return 0;
}
