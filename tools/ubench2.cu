// ubench2.cu -- constant-bank (uniform datapath) tap delivery: LDCU.32/.64/.128 with a dynamic uniform index and
// static indices, feeding FFMA / FFMA2 with UR operands.  Reports complex FMAs (warp-level) per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
struct Taps { float4 t4[1024]; };

__device__ __forceinline__ void cfma(unsigned long long &acc, float t, unsigned long long w) {
    unsigned long long tt;
    asm("mov.b64 %0, {%1,%1};" : "=l"(tt) : "f"(t));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(tt), "l"(w));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const __grid_constant__ Taps P, float *out, const float *in, int dyn) {
    unsigned long long acc[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = 0;
        asm("mov.b64 %0, {%1,%2};" : "=l"(w[i]) : "f"(in[threadIdx.x + i]), "f"(in[threadIdx.x + i + 8]));
    }
    const float *tf = reinterpret_cast<const float *>(P.t4);
    const float2 *t2 = reinterpret_cast<const float2 *>(P.t4);
    if (MODE == 0) {            // dynamic LDCU.32, 1 per complex FMA
        int off = dyn;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) cfma(acc[i & 7], tf[off + i], w[i & 7]);
            off += 16; if (off >= 4000) off -= 4000;
        }
    } else if (MODE == 1) {     // dynamic LDCU.64
        int off = dyn;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { float2 t = t2[off + i]; cfma(acc[(2 * i) & 7], t.x, w[(2 * i) & 7]); cfma(acc[(2 * i + 1) & 7], t.y, w[(2 * i + 1) & 7]); }
            off += 8; if (off >= 2000) off -= 2000;
        }
    } else if (MODE == 2) {     // dynamic LDCU.128
        int off = dyn;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 t = P.t4[off + i];
                cfma(acc[(4 * i) & 7], t.x, w[(4 * i) & 7]); cfma(acc[(4 * i + 1) & 7], t.y, w[(4 * i + 1) & 7]);
                cfma(acc[(4 * i + 2) & 7], t.z, w[(4 * i + 2) & 7]); cfma(acc[(4 * i + 3) & 7], t.w, w[(4 * i + 3) & 7]);
            }
            off += 4; if (off >= 1000) off -= 1000;
        }
    } else if (MODE == 3) {     // static LDCU.128, 256 distinct taps per iteration
        for (int it = 0; it < ITERS / 16; ++it) {
#pragma unroll
            for (int j = 0; j < 256; ++j) cfma(acc[j & 7], tf[j], w[j & 7]);
        }
    } else if (MODE == 4) {     // no tap loads at all (register operand): FFMA2 ceiling with this accumulator pattern
        float t = in[3];
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) cfma(acc[i & 7], t, w[i & 7]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i])); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int blocks_per_sm) {
    int nb = 148 * blocks_per_sm;
    float *out, *in;
    cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    Taps P; for (int i = 0; i < 1024; ++i) P.t4[i] = make_float4(1.f / (i + 1), 0.5f, 0.25f, 0.125f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<nb, 256>>>(P, out, in, 16);
    cudaEventRecord(e0);
    k<MODE><<<nb, 256>>>(P, out, in, 16);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cf = (double)ITERS * 16 * 8 * nb;
    printf("%-44s blocks/SM=%d time=%.3f ms  complexFMA(FFMA2) warp-instr/clk/SM=%.3f (at 1.965 GHz; peak 2.0)  err=%s\n", name, blocks_per_sm, ms,
           cf / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(in);
}
int main() {
    for (int b : {1, 2, 4}) {
        run<0>("dynamic LDCU.32  : FFMA2 = 1:1", b);
        run<1>("dynamic LDCU.64  : FFMA2 = 1:2", b);
        run<2>("dynamic LDCU.128 : FFMA2 = 1:4", b);
        run<3>("static  LDCU.128 : FFMA2 = 1:4", b);
        run<4>("register operand : FFMA2 only", b);
    }
}
