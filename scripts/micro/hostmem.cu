// Host-memory micro-benchmark for the e2e copy path (diagnostic, not product code):
// how long do cudaHostAlloc / cudaHostRegister / parallel pre-faulting take for a
// matrix-sized buffer, and how fast are the copies that follow?
#include <cuda_runtime.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <chrono>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
int main(int argc, char **argv) {
    size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 6.0) * 1000000000ull;
    void *d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    double t0, t1;
    printf("threads %d\n", omp_get_max_threads());
    // 1. cudaHostAlloc
    void *p; t0 = now(); CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault)); t1 = now();
    printf("cudaHostAlloc %.1f ms\n", 1e3 * (t1 - t0));
    t0 = now(); CK(cudaMemcpyAsync(p, d, bytes, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); t1 = now();
    printf("  D2H pinned %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
    t0 = now(); CK(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); t1 = now();
    printf("  H2D pinned %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
    t0 = now(); CK(cudaFreeHost(p)); t1 = now();
    printf("cudaFreeHost %.1f ms\n", 1e3 * (t1 - t0));
    // 2. fresh malloc + register
    for (int rep = 0; rep < 2; ++rep) {
        void *q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (rep == 1) madvise(q, bytes, MADV_HUGEPAGE);
        t0 = now(); CK(cudaHostRegister(q, bytes, cudaHostRegisterDefault)); t1 = now();
        printf("cudaHostRegister fresh (hugepage advise=%d) %.1f ms\n", rep, 1e3 * (t1 - t0));
        t0 = now(); CK(cudaMemcpyAsync(q, d, bytes, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); t1 = now();
        printf("  D2H registered %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
        t0 = now(); CK(cudaHostUnregister(q)); t1 = now();
        printf("  cudaHostUnregister %.1f ms\n", 1e3 * (t1 - t0));
        t0 = now(); CK(cudaHostRegister(q, bytes, cudaHostRegisterReadOnly)); t1 = now();
        printf("  cudaHostRegister touched readonly %.1f ms\n", 1e3 * (t1 - t0));
        CK(cudaHostUnregister(q));
        // chunked register in a pipeline
        t0 = now();
        size_t chunk = 256ull << 20;
        for (size_t o = 0; o < bytes; o += chunk) { size_t n = bytes - o < chunk ? bytes - o : chunk; CK(cudaHostRegister((char *)q + o, n, cudaHostRegisterDefault)); }
        t1 = now(); printf("  chunked register touched (256 MiB) %.1f ms\n", 1e3 * (t1 - t0));
        t0 = now();
        for (size_t o = 0; o < bytes; o += chunk) CK(cudaHostUnregister((char *)q + o));
        t1 = now(); printf("  chunked unregister %.1f ms\n", 1e3 * (t1 - t0));
        munmap(q, bytes);
    }
    // 3. parallel prefault
    for (int rep = 0; rep < 2; ++rep) {
        char *q = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (rep == 1) madvise(q, bytes, MADV_HUGEPAGE);
        t0 = now();
        #pragma omp parallel for schedule(static)
        for (size_t o = 0; o < bytes; o += 4096) q[o] = 0;
        t1 = now(); printf("parallel prefault (hugepage=%d) %.1f ms\n", rep, 1e3 * (t1 - t0));
        t0 = now(); CK(cudaMemcpyAsync(q, d, bytes, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); t1 = now();
        printf("  D2H pageable touched %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
        t0 = now(); CK(cudaMemcpyAsync(d, q, bytes, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); t1 = now();
        printf("  H2D pageable %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
        // staged: threads memcpy chunks into a pinned ring, async H2D per chunk
        const int NB = 4; size_t cb = 64ull << 20; void *ring[NB]; cudaEvent_t ev[NB];
        for (int i = 0; i < NB; ++i) { CK(cudaHostAlloc(&ring[i], cb, cudaHostAllocDefault)); CK(cudaEventCreate(&ev[i])); }
        t0 = now();
        size_t nchunks = (bytes + cb - 1) / cb;
        for (size_t c = 0; c < nchunks; ++c) {
            int b = c % NB; size_t o = c * cb, n = bytes - o < cb ? bytes - o : cb;
            if (c >= NB) CK(cudaEventSynchronize(ev[b]));
            #pragma omp parallel
            { int t = omp_get_thread_num(), T = omp_get_num_threads(); size_t per = (n + T - 1) / T; size_t a = per * t; if (a < n) memcpy((char *)ring[b] + a, q + o + a, a + per <= n ? per : n - a); }
            CK(cudaMemcpyAsync((char *)d + o, ring[b], n, cudaMemcpyHostToDevice, s)); CK(cudaEventRecord(ev[b], s));
        }
        CK(cudaStreamSynchronize(s)); t1 = now();
        printf("  H2D staged via pinned ring %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
        t0 = now();
        for (size_t c = 0; c < nchunks + 1; ++c) {
            if (c < nchunks) { int b = c % NB; size_t o = c * cb, n = bytes - o < cb ? bytes - o : cb;
                CK(cudaMemcpyAsync(ring[b], (char *)d + o, n, cudaMemcpyDeviceToHost, s)); CK(cudaEventRecord(ev[b], s)); }
            if (c >= 1) { size_t cc = c - 1; int b = cc % NB; size_t o = cc * cb, n = bytes - o < cb ? bytes - o : cb;
                CK(cudaEventSynchronize(ev[b]));
                #pragma omp parallel
                { int t = omp_get_thread_num(), T = omp_get_num_threads(); size_t per = (n + T - 1) / T; size_t a = per * t; if (a < n) memcpy(q + o + a, (char *)ring[b] + a, a + per <= n ? per : n - a); } }
        }
        t1 = now();
        printf("  D2H staged via pinned ring %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
        for (int i = 0; i < NB; ++i) cudaFreeHost(ring[i]);
        munmap(q, bytes);
    }
    return 0;
}
