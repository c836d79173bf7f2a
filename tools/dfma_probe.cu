// dfma_probe.cu -- measured FP64 rates on this GPU: DFMA (CUDA cores) and DMMA m8n8k4 (mma.sync), per SM per clock.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_probe tools/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, double a, double b, int iters) {
    double c[8];
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void k_dmma(double *out, double a, double b, int iters) {
    double c[4][2];
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
int main() {
    double *d; cudaMalloc(&d, 8);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int which = 0; which < 2; ++which) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_dfma<<<pr.multiProcessorCount, warps * 32>>>(d, 1.0000001, 1e-9, iters);
                else k_dmma<<<pr.multiProcessorCount, warps * 32>>>(d, 1.0000001, 1e-9, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            const double fma_per_sm = which == 0 ? (double)iters * 8 * warps * 32 : (double)iters * 4 * warps * 256;
            const double tf = fma_per_sm * pr.multiProcessorCount * 2 / (best * 1e-3) / 1e12;
            printf("%s warps/SM %2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at the nominal %d MHz)\n", which ? "DMMA m8n8k4" : "DFMA       ",
                   warps, best, tf, fma_per_sm / (best * 1e-3) / (clk * 1e3), clk / 1000);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
