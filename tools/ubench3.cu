// ubench3.cu -- what does HBM deliver for the TILED kernel's access pattern?  A pure-TMA copy of a channel-major
// [nch][n] complex64 matrix (row pitch 512 KiB) in boxes of ROWS channels x BOXB bytes, each CTA walking a time
// tile the way k_tiled_c64 does.  No arithmetic: the number is the ceiling the access pattern itself allows.
//   mode 0: copy (load box -> store box)   mode 1: read only   mode 2: write only
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}

struct Cfg {
    int rows, boxb;      // box = rows channels x boxb bytes
    int nbox, ahead;     // ring slots, loads in flight
    int tile_boxes;      // boxes per CTA along time
    int mode;
    int xfast;           // 1: blockIdx.x walks time tiles (like k_tiled_c64); 0: blockIdx.x walks channel groups
    int sbox;            // store box bytes per row (mode 0/2); boxb % sbox == 0
};

__global__ void __launch_bounds__(32) k_copy(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const Cfg c) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int box_bytes = c.rows * c.boxb;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + c.nbox * box_bytes);
    const uint32_t base = smem_u32(smem), bar_base = smem_u32(bars);
    if (threadIdx.x != 0) return;
    for (int i = 0; i < c.nbox; ++i) mbar_init(bar_base + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int tt = c.xfast ? blockIdx.x : blockIdx.y, cg = c.xfast ? blockIdx.y : blockIdx.x;
    const int ch0 = cg * c.rows;
    const int f0 = tt * c.tile_boxes * (c.boxb / 4);      // float coordinate of the tile start
    const int fpb = c.boxb / 4;
    const int nstore = c.boxb / c.sbox;
    int issued = 0;
    for (int b = 0; b < c.tile_boxes; ++b) {
        if (c.mode != 2) {
            for (; issued < c.tile_boxes && issued <= b + c.ahead; ++issued) {
                // slot reuse: the store that last read this slot must have finished reading
                if (c.mode == 0 && issued >= c.nbox) {
                    const int allowed = c.nbox - 1 - c.ahead;      // stores that may still be reading their slots
                    if (allowed >= 4) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(4) : "memory");
                    else if (allowed >= 2) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(2) : "memory");
                    else asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(1) : "memory");
                }
                const uint32_t bar = bar_base + 8 * (issued % c.nbox);
                mbar_expect_tx(bar, box_bytes);
                tma_load_2d(base + (issued % c.nbox) * box_bytes, &tmx, f0 + issued * fpb, ch0, bar);
            }
            mbar_wait(bar_base + 8 * (b % c.nbox), (uint32_t)((b / c.nbox) & 1));
        }
        if (c.mode != 1) {
            for (int s = 0; s < nstore; ++s)
                tma_store_2d(&tmy, f0 + b * fpb + s * (c.sbox / 4), ch0, base + (b % c.nbox) * box_bytes + s * c.rows * c.sbox);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (c.mode == 2) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(4) : "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(0) : "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const long long nch = 8192, n = 65536;            // complex64: 512 KiB per channel row, 4 GiB per matrix
    float *x, *y;
    CK(cudaMalloc(&x, nch * n * 8));
    CK(cudaMalloc(&y, nch * n * 8));
    CK(cudaMemset(x, 1, nch * n * 8));
    CK(cudaMemset(y, 0, nch * n * 8));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    CK(cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    // rows, boxb, nbox, ahead, tile bytes per row, mode, xfast, sbox, smem pad (KB; sets CTAs/SM), promo
    struct Run { int rows, boxb, nbox, ahead, tile_row_bytes, mode, xfast, sbox, smem_kb, promo; };
    const Run runs[] = {
        // the tiled kernel's pattern: [64 ch][64 B] loads, ring 8, ~2 ahead, 4 CTAs/SM, [64][32 B] stores
        {64, 64, 8, 2, 8192, 0, 1, 32, 50, 1}, {64, 64, 8, 2, 8192, 1, 1, 32, 50, 1}, {64, 64, 8, 2, 8192, 2, 1, 32, 50, 1},
        {64, 64, 8, 6, 8192, 0, 1, 32, 50, 1}, {64, 64, 8, 6, 8192, 1, 1, 32, 50, 1},
        {64, 64, 8, 6, 8192, 0, 1, 64, 50, 1},
        {64, 64, 16, 12, 8192, 0, 1, 64, 100, 1}, {64, 64, 16, 12, 8192, 1, 1, 64, 100, 1},
        {64, 64, 8, 6, 8192, 0, 1, 64, 33, 1}, {64, 64, 8, 6, 8192, 1, 1, 64, 33, 1},
        {64, 64, 8, 6, 8192, 0, 0, 64, 50, 1}, {64, 64, 8, 6, 8192, 1, 0, 64, 50, 1},
        {64, 64, 8, 6, 8192, 1, 1, 64, 50, 0}, {64, 64, 8, 6, 8192, 1, 1, 64, 50, 2},
        // wider rows, same 4 KiB boxes
        {32, 128, 8, 6, 8192, 0, 1, 128, 50, 1}, {32, 128, 8, 6, 8192, 1, 1, 128, 50, 1}, {32, 128, 8, 6, 8192, 2, 1, 128, 50, 1},
        {16, 256, 8, 6, 8192, 0, 1, 256, 50, 1}, {16, 256, 8, 6, 8192, 1, 1, 256, 50, 1}, {16, 256, 8, 6, 8192, 2, 1, 256, 50, 1},
        {8, 512, 8, 6, 8192, 0, 1, 512, 50, 1}, {8, 512, 8, 6, 8192, 1, 1, 512, 50, 1}, {8, 512, 8, 6, 8192, 2, 1, 512, 50, 1},
        {4, 1024, 8, 6, 8192, 0, 1, 1024, 50, 1}, {4, 1024, 8, 6, 8192, 1, 1, 1024, 50, 1},
        // bigger boxes: [64][128 B] = 8 KiB, [32][256 B]
        {64, 128, 8, 6, 8192, 0, 1, 128, 66, 1}, {64, 128, 8, 6, 8192, 1, 1, 128, 66, 1},
        {64, 128, 4, 3, 8192, 0, 1, 128, 50, 1}, {64, 128, 4, 3, 8192, 1, 1, 128, 50, 1},
        {32, 256, 8, 6, 8192, 0, 1, 256, 66, 1}, {32, 256, 8, 6, 8192, 1, 1, 256, 66, 1},
        // 2 / 1 CTAs per SM with deep rings
        {64, 64, 24, 20, 8192, 0, 1, 64, 110, 1}, {64, 64, 24, 20, 8192, 1, 1, 64, 110, 1},
        {64, 128, 24, 20, 16384, 0, 1, 128, 200, 1}, {64, 128, 24, 20, 16384, 1, 1, 128, 200, 1},
    };
    printf("%-5s %-5s %-4s %-5s %-8s %-4s %-5s %-5s %-7s %-5s | %-8s %-10s\n", "rows", "boxb", "nbox", "ahead", "tileB", "mode", "xfast", "sbox", "smemKB", "promo", "ms", "GB/s");
    for (const Run &r : runs) {
        CUtensorMap tmx, tmy;
        cuuint64_t dims[2] = {(cuuint64_t)(2 * n), (cuuint64_t)nch};
        cuuint64_t strides[1] = {(cuuint64_t)n * 8};
        cuuint32_t es[2] = {1, 1};
        cuuint32_t box[2] = {(cuuint32_t)(r.boxb / 4), (cuuint32_t)r.rows};
        cuuint32_t sbox[2] = {(cuuint32_t)(r.sbox / 4), (cuuint32_t)r.rows};
        const CUtensorMapL2promotion promo = r.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : r.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        if (enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode x failed\n"); continue; }
        if (enc(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, y, dims, strides, sbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode y failed\n"); continue; }
        Cfg c{r.rows, r.boxb, r.nbox, r.ahead, r.tile_row_bytes / r.boxb, r.mode, r.xfast, r.sbox};
        const int ntt = (int)(n * 8 / r.tile_row_bytes), ncg = (int)(nch / r.rows);
        dim3 grid(r.xfast ? ntt : ncg, r.xfast ? ncg : ntt);
        const int smem = r.smem_kb * 1024;
        if (smem < r.nbox * r.rows * r.boxb + 8 * r.nbox) { printf("smem too small\n"); continue; }
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            CK(cudaEventRecord(e0));
            k_copy<<<grid, 32, smem>>>(tmx, tmy, c);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        const double bytes = (double)nch * n * 8 * (r.mode == 0 ? 2 : 1);
        printf("%-5d %-5d %-4d %-5d %-8d %-4d %-5d %-5d %-7d %-5d | %-8.3f %-10.1f\n", r.rows, r.boxb, r.nbox, r.ahead, r.tile_row_bytes, r.mode, r.xfast, r.sbox, r.smem_kb, r.promo, best, bytes / best * 1e-6);
        fflush(stdout);
    }
    return 0;
}
