// Ceilings that bound the EM pass on this GPU (diagnostic, not product code):
// read-only HBM streaming (LDG.128 grid-stride and TMA bulk ring) and the fp64
// FMA issue rate.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void __launch_bounds__(512) ldg_sum(const double2 *__restrict__ p, size_t n, double *out) {
    double a = 0, b = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n; i += 4 * st) {
        double2 v0 = p[i], v1 = p[i + st], v2 = p[i + 2 * st], v3 = p[i + 3 * st];
        a += v0.x + v1.x + v2.x + v3.x; b += v0.y + v1.y + v2.y + v3.y;
    }
    for (; i < n; i += st) { double2 v = p[i]; a += v.x; b += v.y; }
    if (a + b == 12345.678) out[0] = a;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// bulk ring without math: thread 0 issues, everybody waits, one LDS per thread per stage
__global__ void __launch_bounds__(512, 1) bulk_ring(const double *__restrict__ src, size_t row_doubles, size_t n_rows, int n_stages, double *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    double *stages = (double *)smem;
    uint64_t *full = (uint64_t *)(smem + (size_t)n_stages * row_doubles * 8);
    const int tid = threadIdx.x;
    size_t r0 = n_rows * blockIdx.x / gridDim.x, r1 = n_rows * (blockIdx.x + 1) / gridDim.x, n_my = r1 - r0;
    const double *my = src + r0 * row_doubles;
    uint32_t bytes = (uint32_t)(row_doubles * 8);
    if (tid == 0) { for (int s = 0; s < n_stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid == 0) for (size_t q = 0; q < n_my && q < (size_t)n_stages; ++q) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[q])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(stages + q * row_doubles)), "l"(my + q * row_doubles), "r"(bytes), "r"(smem_u32(&full[q])) : "memory");
    }
    double acc = 0;
    for (size_t q = 0; q < n_my; ++q) {
        int s = (int)(q % n_stages); uint32_t par = (uint32_t)((q / n_stages) & 1);
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&full[s])), "r"(par) : "memory");
        acc += stages[(size_t)s * row_doubles + tid];
        __syncthreads();
        if (tid == 0 && q + n_stages < n_my) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(stages + (size_t)s * row_doubles)), "l"(my + (q + n_stages) * row_doubles), "r"(bytes), "r"(smem_u32(&full[s])) : "memory");
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

__global__ void __launch_bounds__(512) dfma_rate(double *out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main() {
    size_t row = 5408, n_rows = 138569, n = row * n_rows;
    double *d, *out; CK(cudaMalloc(&d, n * 8)); CK(cudaMemset(d, 0, n * 8)); CK(cudaMalloc(&out, 8 * 148 * 512 * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int mult : {1, 2, 4, 8, 16}) {
        for (int rep = 0; rep < 3; ++rep) ldg_sum<<<148 * mult, 512>>>((const double2 *)d, n / 2, out);
        CK(cudaEventRecord(e0)); for (int rep = 0; rep < 10; ++rep) ldg_sum<<<148 * mult, 512>>>((const double2 *)d, n / 2, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("ldg_sum grid 148x%-2d : %.3f ms  %.0f GB/s\n", mult, ms / 10, n * 8 / (ms / 10) / 1e6);
    }
    for (int st : {3, 4, 5}) {
        size_t smem = (size_t)st * row * 8 + 64;
        CK(cudaFuncSetAttribute(bulk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int rep = 0; rep < 3; ++rep) bulk_ring<<<148, 512, smem>>>(d, row, n_rows, st, out);
        CK(cudaEventRecord(e0)); for (int rep = 0; rep < 10; ++rep) bulk_ring<<<148, 512, smem>>>(d, row, n_rows, st, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("bulk_ring stages %d      : %.3f ms  %.0f GB/s\n", st, ms / 10, n * 8 / (ms / 10) / 1e6);
    }
    // half-row chunks, 10 stages
    {
        size_t hr = row / 2; int st = 10; size_t smem = (size_t)st * hr * 8 + 128;
        CK(cudaFuncSetAttribute(bulk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int rep = 0; rep < 3; ++rep) bulk_ring<<<148, 512, smem>>>(d, hr, n_rows * 2, st, out);
        CK(cudaEventRecord(e0)); for (int rep = 0; rep < 10; ++rep) bulk_ring<<<148, 512, smem>>>(d, hr, n_rows * 2, st, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("bulk_ring half rows x10   : %.3f ms  %.0f GB/s\n", ms / 10, n * 8 / (ms / 10) / 1e6);
    }
    int iters = 20000;
    dfma_rate<<<148 * 2, 512>>>(out, 100);
    CK(cudaEventRecord(e0)); dfma_rate<<<148 * 2, 512>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fmas = (double)148 * 2 * 512 * 8 * iters;
    printf("dfma: %.3f ms, %.2f TFMA/s = %.1f DFMA/clk/SM at 1.965 GHz\n", ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.965e9);
    CK(cudaGetLastError());
    return 0;
}
