// h2d_ceiling.cu -- the host <-> device copy ceiling of the box, for bench.py's end-to-end figure.
// Plain cudaMemcpyAsync loops, the same bytes per step as the C5 shard (4 GiB in, 3.68 GiB out per GPU), pinned host
// memory, H2D and D2H concurrently on two streams per device, one host thread per device in ONE process.
// Prints, for n = 1, 2, 4, 8 devices (as many as visible): aggregate GB/s each way and combined, and the Gout/s of
// the 147//160 complex64 resampler that bandwidth would allow (16.71 bytes of traffic per output).
// Build: nvcc -O3 -std=c++17 -o h2d_ceiling h2d_ceiling.cu -lpthread      Not part of the library.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Dev {
    void *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
};

int main(int argc, char **argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const size_t in_bytes = argc > 1 ? (size_t)atoll(argv[1]) << 20 : (size_t)1 << 30;       // per copy
    const size_t out_bytes = in_bytes / 160 * 147;
    const int reps = argc > 2 ? atoi(argv[2]) : 4;
    const bool wc = argc > 3 && atoi(argv[3]);                                                // write-combined sources
    printf("devices %d, %zu MiB in / %zu MiB out per copy, %d copies per direction per device, %s\n", ndev, in_bytes >> 20,
           out_bytes >> 20, reps, wc ? "write-combined H2D source" : "plain pinned");
    std::vector<Dev> dv(ndev);
    for (int i = 0; i < ndev; ++i) {
        CK(cudaSetDevice(i));
        CK(cudaMalloc(&dv[i].d_in, in_bytes));
        CK(cudaMalloc(&dv[i].d_out, out_bytes));
        CK(cudaHostAlloc(&dv[i].h_in, in_bytes, cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0)));
        CK(cudaHostAlloc(&dv[i].h_out, out_bytes, cudaHostAllocPortable));
        memset(dv[i].h_in, 1, in_bytes);
        memset(dv[i].h_out, 0, out_bytes);
        CK(cudaStreamCreateWithFlags(&dv[i].s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&dv[i].s_out, cudaStreamNonBlocking));
    }
    for (int mode = 0; mode < 3; ++mode) {          // 0: H2D only, 1: D2H only, 2: both at once
        for (int n = 1; n <= ndev; n *= 2) {
            auto work = [&](int i) {
                CK(cudaSetDevice(i));
                for (int r = 0; r < reps; ++r) {
                    if (mode != 1) CK(cudaMemcpyAsync(dv[i].d_in, dv[i].h_in, in_bytes, cudaMemcpyHostToDevice, dv[i].s_in));
                    if (mode != 0) CK(cudaMemcpyAsync(dv[i].h_out, dv[i].d_out, out_bytes, cudaMemcpyDeviceToHost, dv[i].s_out));
                }
                CK(cudaStreamSynchronize(dv[i].s_in));
                CK(cudaStreamSynchronize(dv[i].s_out));
            };
            for (int i = 0; i < n; ++i) work(i);    // warm-up
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int i = 0; i < n; ++i) th.emplace_back(work, i);
            for (auto &t : th) t.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const double gin = mode != 1 ? (double)n * reps * in_bytes / dt / 1e9 : 0.0;
            const double gout = mode != 0 ? (double)n * reps * out_bytes / dt / 1e9 : 0.0;
            printf("%-8s n=%d: H2D %7.1f GB/s  D2H %7.1f GB/s  total %7.1f GB/s", mode == 0 ? "h2d" : mode == 1 ? "d2h" : "duplex", n, gin, gout, gin + gout);
            if (mode == 2) printf("  -> ceiling %.2f Gout/s for 147//160 complex64 (16.71 B per output)", (gin + gout) / 16.71);
            printf("\n");
        }
    }
    return 0;
}
