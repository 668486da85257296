// Device code of kernel 2 (the EM fit): everything csrc/em.cu launches.  See em.cu for the
// formulation and the host side.  Kept in a header of its own so that tests/emul/ can compile
// the very same kernel bodies for the host (MXB_CPU_EMUL: cooperative fibers stand in for the
// threads of a CTA, a few dozen lines stand in for mbarriers, bulk copies and shuffles) and
// check their index arithmetic and synchronisation without a GPU.  The product never defines
// MXB_CPU_EMUL; nvcc sees exactly the code that used to sit at the top of em.cu.
#pragma once

#ifdef MXB_CPU_EMUL
#define MXB_DYN_SHARED extern          /* the emulator defines the arrays */
#else
#include <cooperative_groups.h>

#include "common.cuh"
#define MXB_DYN_SHARED extern __shared__
#endif

#include <math.h>
#include <stdint.h>

#include <type_traits>

namespace mxb {

// Device-resident control block: lets every kernel of an iteration early-exit
// once the run has converged, so iterations can be enqueued ahead of the host.
struct EmState {
    int done;            // 0 running, 1 converged, 2 max_iter reached
    int cur;             // index of the current (input) proportions buffer
    int bad;             // rows whose mixture likelihood underflowed to 0
    int pad;
    long long iters;
    long long max_iter;
    double tol;
    double delta;
};

constexpr int kPassThreads = 512;
constexpr int kPassWarps = kPassThreads / 32;
constexpr int kPassGroup = 2;       // rows reduced together per block barrier
constexpr int kMaxNC = 8;           // column chunks (double2) per thread
constexpr int kLdAlign = 16;        // row stride of L in doubles (128 B)

// ---- small PTX helpers ------------------------------------------------------
// (tests/emul/ compiles the kernels of this file for the host with MXB_CPU_EMUL and its own
// versions of these helpers: an interleaving model of the CTA, test infrastructure only)
#ifndef MXB_CPU_EMUL
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier.
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes,
                                          uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load_u32(uint32_t dst_smem, const void *src, uint32_t bytes,
                                              uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Programmatic dependent launch: a kernel launched with the stream-serialization attribute
// may start while its predecessor in the stream is still running; it must not touch what the
// predecessor writes before pdl_wait().  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// loads through 32-bit shared-window addresses (record base, warp-uniform, + per-thread offset:
// ptxas folds the sum into the [R + UR] form of LDS)
__device__ __forceinline__ uint2 lds_v2_u32(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_v2_f64(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}

#endif  // MXB_CPU_EMUL

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// Block-wide deterministic reductions (every thread gets the result).
template <int kThreads>
__device__ __forceinline__ double block_sum(double v, double *scratch /*[kThreads/32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // scratch free
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += scratch[i];
    return t;
}
template <int kThreads>
__device__ __forceinline__ double block_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = scratch[0];
#pragma unroll
    for (int i = 1; i < kThreads / 32; ++i) t = fmax(t, scratch[i]);
    return t;
}

// ---- M -> L ------------------------------------------------------------------
// L_ij = exp(M_ij - max_j M_ij); padding columns [n_cols, ld) are 0.
__global__ void __launch_bounds__(256)
to_linear_kernel(const double *__restrict__ m, int64_t n_rows, int64_t n_cols, int64_t ld,
                 double *__restrict__ lin) {
    __shared__ double scratch[8];
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const double *row = m + r * n_cols;
        double mx = -INFINITY;
        bool has_nan = false;
        for (int64_t j = threadIdx.x; j < n_cols; j += 256) {
            double v = row[j];
            has_nan |= (v != v);
            mx = fmax(mx, v);
        }
        mx = block_max<256>(mx, scratch);
        double *dst = lin + r * ld;
        for (int64_t j = threadIdx.x; j < ld; j += 256)
            dst[j] = (j < n_cols) ? exp(row[j] - mx) : 0.0;  // row re-read hits L1/L2
        (void)has_nan;
    }
}

// ---- dictionary-coded rows -----------------------------------------------------
// A row of the matrix built by kernel 1 holds few distinct values (one per class of
// haplotypes with the same match pattern: median 56, at most 256 for 93 % of the config-2
// rows), and so does its row of L.  Such a row is stored losslessly as one byte per cell
// plus a table of 256 doubles -- ld + 2048 bytes instead of 8 ld (5.8x fewer at H = 5408) --
// and the pass kernel looks the values up in shared memory: the same numbers enter the same
// sums in the same order, with a fraction of the HBM traffic.  Rows with more distinct
// values ("dense rows") are gathered into a small fp64 matrix of their own and go through
// the uncoded pass kernel.  Record of row r: [ld code bytes][256 doubles] at r * rec_bytes;
// the record of a dense row has a zeroed table and weight 0, so it adds exactly nothing.
constexpr int kDictSize = 256;
constexpr int kDictSlots = 1024;      // hash slots of the coder (at most 512 ever taken)
constexpr int kPackThreads = 256;
constexpr unsigned long long kDictEmpty = 0xFFFFFFFFFFFFFFFFull;  // a NaN L never holds

__global__ void __launch_bounds__(kPackThreads)
em_pack_kernel(const double *__restrict__ lin, int64_t n_rows, int64_t ld,
               const double *__restrict__ weights, unsigned char *__restrict__ rec,
               int64_t rec_bytes, int *__restrict__ dense_flag, double *__restrict__ w_coded) {
    __shared__ unsigned long long keys[kDictSlots];
    __shared__ unsigned short ids[kDictSlots];
    __shared__ int count;
    MXB_DYN_SHARED unsigned short cell_slot[];   // [ld] hash slot of every cell
    const int tid = threadIdx.x;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        for (int i = tid; i < kDictSlots; i += kPackThreads) keys[i] = kDictEmpty;
        if (tid == 0) count = 0;
        __syncthreads();
        const double *src = lin + r * ld;
        for (int64_t j = tid; j < ld; j += kPackThreads) {
            if (*reinterpret_cast<volatile int *>(&count) > kDictSize) break;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(src[j]);
            if (bits == kDictEmpty) { atomicAdd(&count, kDictSize + 1); break; }
            unsigned h = (unsigned)((bits * 0x9E3779B97F4A7C15ull) >> 54);
            while (true) {
                const unsigned long long old = atomicCAS(&keys[h], kDictEmpty, bits);
                if (old == kDictEmpty) {   // first sight of the value: next free code
                    ids[h] = (unsigned short)atomicAdd(&count, 1);
                    break;
                }
                if (old == bits) break;
                h = (h + 1) & (kDictSlots - 1);
            }
            cell_slot[j] = (unsigned short)h;
        }
        __syncthreads();
        const bool coded = count <= kDictSize;   // block-uniform
        unsigned char *out = rec + r * rec_bytes;
        double *tab = reinterpret_cast<double *>(out + ld);
        for (int i = tid; i < kDictSize; i += kPackThreads) tab[i] = 0.0;
        __syncthreads();
        if (coded) {
            for (int64_t j = tid; j < ld; j += kPackThreads) out[j] = (unsigned char)ids[cell_slot[j]];
            for (int i = tid; i < kDictSlots; i += kPackThreads)
                if (keys[i] != kDictEmpty) tab[ids[i]] = __longlong_as_double((long long)keys[i]);
        } else {
            for (int64_t j = tid; j < ld; j += kPackThreads) out[j] = 0;
        }
        if (tid == 0) {
            dense_flag[r] = coded ? 0 : 1;
            w_coded[r] = coded ? weights[r] : 0.0;
        }
        __syncthreads();
    }
}

// Experimental (MXB_EM_CODED_PAIRS=1, not yet run on a GPU): the dictionary holds the values
// of a *chunk* -- the two adjacent cells 2c, 2c + 1 one pass-kernel thread handles together --
// instead of single cells.  91.6 % of the config-2 rows have at most 256 distinct chunks
// (92.4 % have at most 256 distinct cells: tests/analysis/pair_codes.py), so about the same
// rows stay coded, and a chunk costs the pass one table lookup (LDS.128) instead of two
// (LDS.64) and half the index arithmetic.  Record of row r, pair_rec_bytes(T) bytes:
//   [T x 8 code bytes: byte k of thread t = code of chunk t + k * T]   (T = threads of the pass)
//   [256 x double2: the two values of a chunk]
// so a thread fetches all its codes of a row with one 8-byte load.  The hash key of a chunk
// is a 64-bit mix of its two values; every chunk is compared with the chunk that claimed its
// slot afterwards, and a row with a key collision between different chunks simply stays dense.
constexpr int kPairTableBytes = kDictSize * 16;
// record bytes for a pass kernel of `pass_threads` threads: 8 code bytes per thread + the table
__host__ __device__ constexpr int pair_rec_bytes(int pass_threads) {
    return pass_threads * 8 + kPairTableBytes;
}

__device__ __forceinline__ unsigned long long pair_key(unsigned long long a, unsigned long long b) {
    unsigned long long k = (a ^ (b << 29 | b >> 35)) * 0x9E3779B97F4A7C15ull;
    k ^= b * 0xC2B2AE3D27D4EB4Full;
    k ^= k >> 31;
    return k == kDictEmpty ? 0x5851F42D4C957F2Dull : k;
}

__global__ void __launch_bounds__(kPackThreads)
em_pack_pairs_kernel(const double *__restrict__ lin, int64_t n_rows, int64_t ld,
                     const double *__restrict__ weights, unsigned char *__restrict__ rec,
                     int pass_threads, int *__restrict__ dense_flag, double *__restrict__ w_coded) {
    const int rec_bytes = pair_rec_bytes(pass_threads);
    __shared__ unsigned long long keys[kDictSlots];
    __shared__ int rep[kDictSlots];            // the chunk that claimed the slot
    __shared__ unsigned short ids[kDictSlots];
    __shared__ int count;
    __shared__ int clash;
    MXB_DYN_SHARED unsigned short chunk_slot[];   // [ld / 2] hash slot of every chunk
    const int tid = threadIdx.x;
    const int n_chunks = (int)(ld >> 1);
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        for (int i = tid; i < kDictSlots; i += kPackThreads) keys[i] = kDictEmpty;
        if (tid == 0) { count = 0; clash = 0; }
        __syncthreads();
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(lin + r * ld);
        for (int c = tid; c < n_chunks; c += kPackThreads) {
            if (*reinterpret_cast<volatile int *>(&count) > kDictSize) break;
            const ulonglong2 ab = src[c];
            const unsigned long long key = pair_key(ab.x, ab.y);
            unsigned h = (unsigned)(key >> 54);
            while (true) {
                const unsigned long long old = atomicCAS(&keys[h], kDictEmpty, key);
                if (old == kDictEmpty) {   // first sight of the key: next free code
                    rep[h] = c;
                    ids[h] = (unsigned short)atomicAdd(&count, 1);
                    break;
                }
                if (old == key) break;
                h = (h + 1) & (kDictSlots - 1);
            }
            chunk_slot[c] = (unsigned short)h;
        }
        __syncthreads();
        bool coded = count <= kDictSize;   // block-uniform
        if (coded) {
            for (int c = tid; c < n_chunks; c += kPackThreads) {
                const ulonglong2 ab = src[c], rp = src[rep[chunk_slot[c]]];
                if (ab.x != rp.x || ab.y != rp.y) clash = 1;
            }
        }
        __syncthreads();
        coded = coded && clash == 0;
        unsigned char *out = rec + r * (int64_t)rec_bytes;
        for (int i = tid; i < rec_bytes / 8; i += kPackThreads)
            reinterpret_cast<unsigned long long *>(out)[i] = 0ull;
        __syncthreads();
        if (coded) {
            for (int c = tid; c < n_chunks; c += kPackThreads)
                out[(c % pass_threads) * 8 + c / pass_threads] = (unsigned char)ids[chunk_slot[c]];
            ulonglong2 *tab = reinterpret_cast<ulonglong2 *>(out + pass_threads * 8);
            for (int i = tid; i < kDictSlots; i += kPackThreads)
                if (keys[i] != kDictEmpty) tab[ids[i]] = src[rep[i]];
        }
        if (tid == 0) {
            dense_flag[r] = coded ? 0 : 1;
            w_coded[r] = coded ? weights[r] : 0.0;
        }
        __syncthreads();
    }
}

// Positions of the dense rows, in row order (one block: deterministic, N / 1024 steps).
__global__ void __launch_bounds__(1024)
em_dense_list_kernel(const int *__restrict__ dense_flag, int64_t n_rows,
                     int64_t *__restrict__ dense_rows, int64_t *__restrict__ n_dense) {
    __shared__ int warp_tot[32];
    __shared__ int64_t base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int64_t r0 = 0; r0 < n_rows; r0 += 1024) {
        const int64_t r = r0 + tid;
        const int f = (r < n_rows) ? dense_flag[r] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f != 0);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[w];
            if (w < warp) before += t;
            total += t;
        }
        if (f) dense_rows[base + before + __popc(m & ((1u << lane) - 1u))] = r;
        __syncthreads();
        if (tid == 0) base += total;
        __syncthreads();
    }
    if (tid == 0) *n_dense = base;
}

// dst[i] = lin[dense_rows[i]], w_dst[i] = weights[dense_rows[i]]
__global__ void __launch_bounds__(256)
em_gather_rows_kernel(const double *__restrict__ lin, int64_t ld, const double *__restrict__ weights,
                      const int64_t *__restrict__ dense_rows, int64_t n_dense,
                      double *__restrict__ dst, double *__restrict__ w_dst) {
    for (int64_t i = blockIdx.x; i < n_dense; i += gridDim.x) {
        const int64_t r = dense_rows[i];
        const double2 *src = reinterpret_cast<const double2 *>(lin + r * ld);
        double2 *d = reinterpret_cast<double2 *>(dst + i * ld);
        for (int64_t j = threadIdx.x; j < ld / 2; j += 256) d[j] = src[j];
        if (threadIdx.x == 0) w_dst[i] = weights[r];
    }
}

// MXB_EM_CODED_COMPACT=1 (experimental): the coded pass skips the dense rows instead of
// running over their empty records -- records and weights of the coded rows only, in row order.
__global__ void em_flag_invert_kernel(int *__restrict__ flag, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = flag[i] ? 0 : 1;
}
__global__ void __launch_bounds__(256)
em_gather_records_kernel(const unsigned char *__restrict__ rec, int64_t rec_bytes,
                         const double *__restrict__ w_src, const int64_t *__restrict__ rows,
                         int64_t n_out, unsigned char *__restrict__ dst, double *__restrict__ w_dst) {
    for (int64_t i = blockIdx.x; i < n_out; i += gridDim.x) {
        const int64_t r = rows[i];
        const uint4 *src = reinterpret_cast<const uint4 *>(rec + r * rec_bytes);
        uint4 *d = reinterpret_cast<uint4 *>(dst + i * rec_bytes);
        for (int64_t j = threadIdx.x; j < rec_bytes / 16; j += 256) d[j] = src[j];
        if (threadIdx.x == 0) w_dst[i] = w_src[r];
    }
}

// ---- fused E+M pass (fast path) ----------------------------------------------
// Rows are handled two at a time between block barriers.  The two partial dot
// products of a thread are reduced together: the first shuffle step hands row 0
// to lanes 0-15 and row 1 to lanes 16-31, so one 5-step butterfly serves both
// rows; lanes 0 and 16 publish the warp totals, and after the barrier every
// warp folds the 16 + 16 warp totals with a 4-step butterfly over its two
// half-warps, divides once per row and broadcasts the two coefficients.  All
// index arithmetic in the loop is 32-bit and incremental (no 64-bit division).
__device__ __forceinline__ double shfl_xor_f64(double v, int off) {
    return __shfl_xor_sync(0xffffffffu, v, off);
}

// accumulate != 0 adds the column sums to what the launch before this one left in `partials`.
template <int NC>
__global__ void __launch_bounds__(kPassThreads, 1)
em_pass_fast_kernel(const unsigned char *__restrict__ rows, uint32_t row_bytes, int64_t ld,
                    int64_t n_rows, const double *__restrict__ weights,
                    const double *__restrict__ pi0, const double *__restrict__ pi1,
                    EmState *__restrict__ st, double *__restrict__ partials, int n_stages,
                    int accumulate) {
    static_assert(kPassGroup == 2 && kPassWarps == 16, "reduction layout below");
    pdl_launch_dependents();  // the tail kernel may be scheduled as SMs drain

    MXB_DYN_SHARED __align__(128) unsigned char smem_raw[];
    double *scratch = reinterpret_cast<double *>(
        smem_raw + (((size_t)n_stages * row_bytes + 127) & ~(size_t)127));
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 2 * kPassWarps * kPassGroup);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t r_begin = n_rows * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = n_rows * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(r_end - r_begin);
    const unsigned char *my_rows = rows + (size_t)r_begin * row_bytes;
    const double *my_w = weights + r_begin;
    const uint32_t stages_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(full);

    // Prologue: L and the weights do not depend on the previous iteration's tail, so the
    // ring is primed before waiting for it (the loads overlap the tail kernel).
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {
        for (int q = 0; q < n_my && q < n_stages; ++q) {
            mbar_expect_tx(&full[q], row_bytes);
            bulk_load(smem_raw + (size_t)q * row_bytes, my_rows + (size_t)q * row_bytes, row_bytes,
                      &full[q]);
        }
    }
    pdl_wait();  // proportions and control block of the previous iteration are final
    if (st->done) {
        // finished run: the primed loads must land before this CTA's shared memory is released
        for (int q = 0; q < n_my && q < n_stages; ++q) mbar_wait_u32(full_u32 + 8u * (uint32_t)q, 0u);
        return;
    }
    const double *__restrict__ pi = st->cur ? pi1 : pi0;

    // Thread-private column slice: chunk c = tid + k*512 covers doubles 2c, 2c+1.
    const int n_chunks = (int)(ld >> 1);
    double2 pr[NC], tr[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * kPassThreads;
        pr[k] = (c < n_chunks) ? reinterpret_cast<const double2 *>(pi)[c] : make_double2(0.0, 0.0);
        tr[k] = make_double2(0.0, 0.0);
    }
    const bool last_live = tid + (NC - 1) * kPassThreads < n_chunks;  // only chunk NC-1 can be ragged

    int stage = 0;          // ring slot of row q0
    uint32_t phase = 0;     // its mbarrier parity
    int sbuf = 0;
    int bad = 0;
    const bool upper = lane >= 16;
    for (int q0 = 0; q0 < n_my; q0 += kPassGroup) {
        double2 lv[kPassGroup][NC];
        double dot[kPassGroup];
        const int q_mine = q0 + (upper ? 1 : 0);
        const double w_mine = (q_mine < n_my) ? my_w[q_mine] : 0.0;
        int s_of[kPassGroup];
        int s = stage;
        uint32_t ph = phase;
#pragma unroll
        for (int g = 0; g < kPassGroup; ++g) {
            s_of[g] = s;
            double dx = 0.0, dy = 0.0;
            if (q0 + g < n_my) {
                mbar_wait_u32(full_u32 + 8u * (uint32_t)s, ph);
                const double2 *srow = reinterpret_cast<const double2 *>(
                    smem_raw + (size_t)s * row_bytes) + tid;
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    if (k < NC - 1 || last_live) lv[g][k] = srow[k * kPassThreads];
                    else lv[g][k] = make_double2(0.0, 0.0);
                    dx = fma(lv[g][k].x, pr[k].x, dx);
                    dy = fma(lv[g][k].y, pr[k].y, dy);
                }
            } else {
#pragma unroll
                for (int k = 0; k < NC; ++k) lv[g][k] = make_double2(0.0, 0.0);
            }
            dot[g] = dx + dy;
            if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        // both rows in one butterfly: lanes 0-15 end up with row 0, lanes 16-31 with row 1
        double v = (upper ? dot[1] : dot[0]) + shfl_xor_f64(upper ? dot[0] : dot[1], 16);
        v += shfl_xor_f64(v, 8);
        v += shfl_xor_f64(v, 4);
        v += shfl_xor_f64(v, 2);
        v += shfl_xor_f64(v, 1);
        double *sc = scratch + sbuf * (kPassWarps * kPassGroup);
        if ((lane & 15) == 0) sc[(lane >> 4) * kPassWarps + warp] = v;
        __syncthreads();  // all reads of this group's stages are done; warp totals visible
        if (tid == 0) {
#pragma unroll
            for (int g = 0; g < kPassGroup; ++g) {
                const int q = q0 + g + n_stages;
                if (q < n_my) {
                    const uint32_t bar = full_u32 + 8u * (uint32_t)s_of[g];
                    mbar_expect_tx_u32(bar, row_bytes);
                    bulk_load_u32(stages_u32 + (uint32_t)s_of[g] * row_bytes,
                                  my_rows + (size_t)q * row_bytes, row_bytes, bar);
                }
            }
        }
        // 16 warp totals per row sit in sc[0..15] / sc[16..31]: one value per lane
        double t = sc[lane];
        t += shfl_xor_f64(t, 8);
        t += shfl_xor_f64(t, 4);
        t += shfl_xor_f64(t, 2);
        t += shfl_xor_f64(t, 1);
        double coef_mine = 0.0;
        if (w_mine != 0.0) {
            coef_mine = w_mine / t;
            bad |= (t == 0.0);
        }
        const double coef0 = __shfl_sync(0xffffffffu, coef_mine, 0);
        const double coef1 = __shfl_sync(0xffffffffu, coef_mine, 16);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            tr[k].x = fma(coef0, lv[0][k].x, tr[k].x);
            tr[k].y = fma(coef0, lv[0][k].y, tr[k].y);
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            tr[k].x = fma(coef1, lv[1][k].x, tr[k].x);
            tr[k].y = fma(coef1, lv[1][k].y, tr[k].y);
        }
        stage = s;
        phase = ph;
        sbuf ^= 1;
    }

    double2 *out = reinterpret_cast<double2 *>(partials + (size_t)blockIdx.x * ld);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * kPassThreads;
        if (c < n_chunks) {
            if (accumulate) {
                const double2 prev = out[c];
                out[c] = make_double2(prev.x + tr[k].x, prev.y + tr[k].y);
            } else {
                out[c] = tr[k];
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0 && warp == 0) atomicAdd(&st->bad, 1);
}

// ---- fused E+M pass over dictionary-coded rows ----------------------------------------
// em_pass_fast_kernel over the records of em_pack_kernel ([ld code bytes][256 doubles], a
// cell is table[code]): same ring, same column slices (chunk c = tid + k*THREADS covers cells
// 2c, 2c+1, one 16-bit load brings both codes), same reduction; only the two values of a
// chunk come from the row's table in shared memory instead of the stage itself.  A record
// is 5.8x smaller than the fp64 row, so this kernel is bound by instruction issue (lookup
// index arithmetic, butterflies, the division) and not by HBM: 0.50 ms for the 138 569
// records of config 2 against 0.87 ms for the fp64 rows (more threads per CTA, eight rows
// per barrier with the values looked up twice, and run-length aware lookups over
// consecutive cells were all measured slower).  The loop runs over full row pairs without
// "is there a second row" tests or zero fills, the odd last row is peeled off, and both
// records are waited for before the lookups of either start: 274 warp instructions per row
// pair in SASS against 307 for the first version of the loop, which kept those tests inside
// (0.568 ms per pass against 0.609 ms, profiles/r1k against r1j).  accumulate != 0 adds the
// column sums to what the launch before this one (the fp64 pass over the dense rows) left
// in `partials`.  THREADS = 384 is the experimental MXB_EM_CODED_T384 shape.
template <int NC, int THREADS = kPassThreads>
__global__ void __launch_bounds__(THREADS, 1)
em_pass_coded_kernel(const unsigned char *__restrict__ rows, uint32_t row_bytes, int64_t ld,
                    int64_t n_rows, const double *__restrict__ weights,
                    const double *__restrict__ pi0, const double *__restrict__ pi1,
                    EmState *__restrict__ st, double *__restrict__ partials, int n_stages,
                    int accumulate) {
    static_assert(kPassGroup == 2 && THREADS % 32 == 0 && THREADS <= kPassThreads,
                  "reduction layout below: at most 16 warp totals per row");
    pdl_launch_dependents();  // the tail kernel may be scheduled as SMs drain

    MXB_DYN_SHARED __align__(128) unsigned char smem_raw[];
    double *scratch = reinterpret_cast<double *>(
        smem_raw + (((size_t)n_stages * row_bytes + 127) & ~(size_t)127));
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 2 * kPassWarps * kPassGroup);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t r_begin = n_rows * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = n_rows * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(r_end - r_begin);
    const unsigned char *my_rows = rows + (size_t)r_begin * row_bytes;
    const double *my_w = weights + r_begin;
    const uint32_t stages_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(full);

    // Prologue: L and the weights do not depend on the previous iteration's tail, so the
    // ring is primed before waiting for it (the loads overlap the tail kernel).
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {
        for (int q = 0; q < n_my && q < n_stages; ++q) {
            mbar_expect_tx(&full[q], row_bytes);
            bulk_load(smem_raw + (size_t)q * row_bytes, my_rows + (size_t)q * row_bytes, row_bytes,
                      &full[q]);
        }
    }
    // totals slots of the warps a smaller CTA does not have (read by the 16-lane butterfly)
    if (THREADS < kPassThreads && tid < 2 * kPassWarps * kPassGroup && (tid & 15) >= THREADS / 32)
        scratch[tid] = 0.0;
    pdl_wait();  // proportions and control block of the previous iteration are final
    if (st->done) {
        // finished run: the primed loads must land before this CTA's shared memory is released
        for (int q = 0; q < n_my && q < n_stages; ++q) mbar_wait_u32(full_u32 + 8u * (uint32_t)q, 0u);
        return;
    }
    const double *__restrict__ pi = st->cur ? pi1 : pi0;

    // Thread-private column slice: chunk c = tid + k*512 covers doubles 2c, 2c+1.
    const int n_chunks = (int)(ld >> 1);
    double2 pr[NC], tr[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * THREADS;
        pr[k] = (c < n_chunks) ? reinterpret_cast<const double2 *>(pi)[c] : make_double2(0.0, 0.0);
        tr[k] = make_double2(0.0, 0.0);
    }
    const bool last_live = tid + (NC - 1) * THREADS < n_chunks;  // only chunk NC-1 can be ragged

    int stage = 0;          // ring slot of row q0
    uint32_t phase = 0;     // its mbarrier parity
    int sbuf = 0;
    int bad = 0;
    const bool upper = lane >= 16;

    // One group of rows between two block barriers: rows q0 and q0 + 1 (kBoth), or the last
    // row alone when the CTA's row count is odd -- the loop over full pairs carries no
    // "is there a second row" tests and no zero fills.  Both records are waited for before the
    // lookups of either start, so the 4 NC table lookups of a thread are independent work.
    auto step = [&](auto both_tag, const int q0) {
        constexpr bool kBoth = decltype(both_tag)::value;
        constexpr int G = kBoth ? 2 : 1;
        double2 lv[G][NC];
        double dot[G];
        int s_of[G];
        const double w_mine = (kBoth || !upper) ? my_w[q0 + (upper ? 1 : 0)] : 0.0;
        int s = stage;
        uint32_t ph = phase;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            s_of[g] = s;
            mbar_wait_u32(full_u32 + 8u * (uint32_t)s, ph);
            if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const unsigned char *srec = smem_raw + (size_t)s_of[g] * row_bytes;
            const uint16_t *codes = reinterpret_cast<const uint16_t *>(srec) + tid;
            const double *tab = reinterpret_cast<const double *>(srec + ld);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                if (k < NC - 1 || last_live) {
                    const unsigned cc = codes[k * THREADS];   // cells 2c, 2c + 1
                    lv[g][k] = make_double2(tab[cc & 0xFFu], tab[cc >> 8]);
                } else {
                    lv[g][k] = make_double2(0.0, 0.0);
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            double dx = 0.0, dy = 0.0;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                dx = fma(lv[g][k].x, pr[k].x, dx);
                dy = fma(lv[g][k].y, pr[k].y, dy);
            }
            dot[g] = dx + dy;
        }
        // both rows in one butterfly: lanes 0-15 end up with row 0, lanes 16-31 with row 1
        const double dot1 = kBoth ? dot[G - 1] : 0.0;
        double v = (upper ? dot1 : dot[0]) + shfl_xor_f64(upper ? dot[0] : dot1, 16);
        v += shfl_xor_f64(v, 8);
        v += shfl_xor_f64(v, 4);
        v += shfl_xor_f64(v, 2);
        v += shfl_xor_f64(v, 1);
        double *sc = scratch + sbuf * (kPassWarps * kPassGroup);
        if ((lane & 15) == 0) sc[(lane >> 4) * kPassWarps + warp] = v;
        __syncthreads();  // all reads of this group's stages are done; warp totals visible
        if (tid == 0) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int q = q0 + g + n_stages;
                if (q < n_my) {
                    const uint32_t bar = full_u32 + 8u * (uint32_t)s_of[g];
                    mbar_expect_tx_u32(bar, row_bytes);
                    bulk_load_u32(stages_u32 + (uint32_t)s_of[g] * row_bytes,
                                  my_rows + (size_t)q * row_bytes, row_bytes, bar);
                }
            }
        }
        // 16 warp totals per row sit in sc[0..15] / sc[16..31]: one value per lane
        double t = sc[lane];
        t += shfl_xor_f64(t, 8);
        t += shfl_xor_f64(t, 4);
        t += shfl_xor_f64(t, 2);
        t += shfl_xor_f64(t, 1);
        double coef_mine = 0.0;
        if (w_mine != 0.0) {
            coef_mine = w_mine / t;
            bad |= (t == 0.0);
        }
        const double coef0 = __shfl_sync(0xffffffffu, coef_mine, 0);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            tr[k].x = fma(coef0, lv[0][k].x, tr[k].x);
            tr[k].y = fma(coef0, lv[0][k].y, tr[k].y);
        }
        if (kBoth) {
            const double coef1 = __shfl_sync(0xffffffffu, coef_mine, 16);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                tr[k].x = fma(coef1, lv[G - 1][k].x, tr[k].x);
                tr[k].y = fma(coef1, lv[G - 1][k].y, tr[k].y);
            }
        }
        stage = s;
        phase = ph;
        sbuf ^= 1;
    };
    int q0 = 0;
    for (; q0 + 1 < n_my; q0 += kPassGroup) step(std::true_type{}, q0);
    if (q0 < n_my) step(std::false_type{}, q0);

    double2 *out = reinterpret_cast<double2 *>(partials + (size_t)blockIdx.x * ld);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * THREADS;
        if (c < n_chunks) {
            if (accumulate) {
                const double2 prev = out[c];
                out[c] = make_double2(prev.x + tr[k].x, prev.y + tr[k].y);
            } else {
                out[c] = tr[k];
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0 && warp == 0) atomicAdd(&st->bad, 1);
}

// em_pass_coded_kernel over chunk-coded records (em_pack_pairs_kernel): one 8-byte load
// brings a thread's codes of a row, one 16-byte lookup the two values of a chunk.
// Experimental, MXB_EM_CODED_PAIRS=1.
template <int NC, int THREADS = kPassThreads>
__global__ void __launch_bounds__(THREADS, 1)
em_pass_coded_pairs_kernel(const unsigned char *__restrict__ rows, uint32_t row_bytes, int64_t ld,
                    int64_t n_rows, const double *__restrict__ weights,
                    const double *__restrict__ pi0, const double *__restrict__ pi1,
                    EmState *__restrict__ st, double *__restrict__ partials, int n_stages,
                    int accumulate) {
    static_assert(kPassGroup == 2 && THREADS % 32 == 0 && THREADS <= kPassThreads && NC <= 8,
                  "at most 16 warp totals per row; a thread's codes of a row fit one 8-byte word");
    pdl_launch_dependents();  // the tail kernel may be scheduled as SMs drain

    MXB_DYN_SHARED __align__(128) unsigned char smem_raw[];
    double *scratch = reinterpret_cast<double *>(
        smem_raw + (((size_t)n_stages * row_bytes + 127) & ~(size_t)127));
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 2 * kPassWarps * kPassGroup);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t r_begin = n_rows * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = n_rows * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(r_end - r_begin);
    const unsigned char *my_rows = rows + (size_t)r_begin * row_bytes;
    const double *my_w = weights + r_begin;
    const uint32_t stages_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(full);

    // Prologue: L and the weights do not depend on the previous iteration's tail, so the
    // ring is primed before waiting for it (the loads overlap the tail kernel).
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {
        for (int q = 0; q < n_my && q < n_stages; ++q) {
            mbar_expect_tx(&full[q], row_bytes);
            bulk_load(smem_raw + (size_t)q * row_bytes, my_rows + (size_t)q * row_bytes, row_bytes,
                      &full[q]);
        }
    }
    // totals slots of the warps a smaller CTA does not have (read by the 16-lane butterfly)
    if (THREADS < kPassThreads && tid < 2 * kPassWarps * kPassGroup && (tid & 15) >= THREADS / 32)
        scratch[tid] = 0.0;
    pdl_wait();  // proportions and control block of the previous iteration are final
    if (st->done) {
        // finished run: the primed loads must land before this CTA's shared memory is released
        for (int q = 0; q < n_my && q < n_stages; ++q) mbar_wait_u32(full_u32 + 8u * (uint32_t)q, 0u);
        return;
    }
    const double *__restrict__ pi = st->cur ? pi1 : pi0;

    // Thread-private column slice: chunk c = tid + k*512 covers doubles 2c, 2c+1.
    const int n_chunks = (int)(ld >> 1);
    double2 pr[NC], tr[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * THREADS;
        pr[k] = (c < n_chunks) ? reinterpret_cast<const double2 *>(pi)[c] : make_double2(0.0, 0.0);
        tr[k] = make_double2(0.0, 0.0);
    }

    int stage = 0;          // ring slot of row q0
    uint32_t phase = 0;     // its mbarrier parity
    int sbuf = 0;
    int bad = 0;
    const bool upper = lane >= 16;

    // One group of rows between two block barriers: rows q0 and q0 + 1 (kBoth), or the last
    // row alone when the CTA's row count is odd -- the loop over full pairs carries no
    // "is there a second row" tests and no zero fills.  Both records are waited for before the
    // lookups of either start, so the 4 NC table lookups of a thread are independent work.
    auto step = [&](auto both_tag, const int q0) {
        constexpr bool kBoth = decltype(both_tag)::value;
        constexpr int G = kBoth ? 2 : 1;
        double2 lv[G][NC];
        double dot[G];
        int s_of[G];
        const double w_mine = (kBoth || !upper) ? my_w[q0 + (upper ? 1 : 0)] : 0.0;
        int s = stage;
        uint32_t ph = phase;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            s_of[g] = s;
            mbar_wait_u32(full_u32 + 8u * (uint32_t)s, ph);
            if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            // 32-bit shared-window addresses: record base (warp-uniform) + per-thread offset
            const uint32_t rec_u32 = stages_u32 + (uint32_t)s_of[g] * row_bytes;
            uint2 cw;   // this thread's codes of the row
            cw = lds_v2_u32(rec_u32 + (uint32_t)tid * 8u);
            const uint32_t tab_u32 = rec_u32 + (uint32_t)(THREADS * 8);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                // a chunk past the end of the row has code 0 and proportion 0: whatever the
                // table holds there adds nothing to the dot product, and its column sum is
                // never written
                const unsigned word = (k < 4) ? cw.x : cw.y;
                const unsigned off = ((word >> (8 * (k & 3))) & 0xFFu) << 4;
                lv[g][k] = lds_v2_f64(tab_u32 + off);
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            double dx = 0.0, dy = 0.0;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                dx = fma(lv[g][k].x, pr[k].x, dx);
                dy = fma(lv[g][k].y, pr[k].y, dy);
            }
            dot[g] = dx + dy;
        }
        // both rows in one butterfly: lanes 0-15 end up with row 0, lanes 16-31 with row 1
        const double dot1 = kBoth ? dot[G - 1] : 0.0;
        double v = (upper ? dot1 : dot[0]) + shfl_xor_f64(upper ? dot[0] : dot1, 16);
        v += shfl_xor_f64(v, 8);
        v += shfl_xor_f64(v, 4);
        v += shfl_xor_f64(v, 2);
        v += shfl_xor_f64(v, 1);
        double *sc = scratch + sbuf * (kPassWarps * kPassGroup);
        if ((lane & 15) == 0) sc[(lane >> 4) * kPassWarps + warp] = v;
        __syncthreads();  // all reads of this group's stages are done; warp totals visible
        if (tid == 0) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int q = q0 + g + n_stages;
                if (q < n_my) {
                    const uint32_t bar = full_u32 + 8u * (uint32_t)s_of[g];
                    mbar_expect_tx_u32(bar, row_bytes);
                    bulk_load_u32(stages_u32 + (uint32_t)s_of[g] * row_bytes,
                                  my_rows + (size_t)q * row_bytes, row_bytes, bar);
                }
            }
        }
        // 16 warp totals per row sit in sc[0..15] / sc[16..31]: one value per lane
        double t = sc[lane];
        t += shfl_xor_f64(t, 8);
        t += shfl_xor_f64(t, 4);
        t += shfl_xor_f64(t, 2);
        t += shfl_xor_f64(t, 1);
        double coef_mine = 0.0;
        if (w_mine != 0.0) {
            coef_mine = w_mine / t;
            bad |= (t == 0.0);
        }
        const double coef0 = __shfl_sync(0xffffffffu, coef_mine, 0);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            tr[k].x = fma(coef0, lv[0][k].x, tr[k].x);
            tr[k].y = fma(coef0, lv[0][k].y, tr[k].y);
        }
        if (kBoth) {
            const double coef1 = __shfl_sync(0xffffffffu, coef_mine, 16);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                tr[k].x = fma(coef1, lv[G - 1][k].x, tr[k].x);
                tr[k].y = fma(coef1, lv[G - 1][k].y, tr[k].y);
            }
        }
        stage = s;
        phase = ph;
        sbuf ^= 1;
    };
    int q0 = 0;
    for (; q0 + 1 < n_my; q0 += kPassGroup) step(std::true_type{}, q0);
    if (q0 < n_my) step(std::false_type{}, q0);

    double2 *out = reinterpret_cast<double2 *>(partials + (size_t)blockIdx.x * ld);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * THREADS;
        if (c < n_chunks) {
            if (accumulate) {
                const double2 prev = out[c];
                out[c] = make_double2(prev.x + tr[k].x, prev.y + tr[k].y);
            } else {
                out[c] = tr[k];
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0 && warp == 0) atomicAdd(&st->bad, 1);
}

// Third version (experimental, MXB_EM_CODED_V3=1, not yet run on a GPU): the same sums in
// the same order, one row at a time, software-pipelined and without a block barrier.
//
// The two-row loop above runs its 16 warps through the same phases in lock step: all of them
// look values up (the LSU is saturated, ~960 of the ~2100 cycles a row pair takes), then all of
// them sit in the butterfly / division chains (nothing to issue), then all of them add.  Here a
// warp publishes its partial dot product of row r + 1 with an mbarrier *arrive* (non-blocking)
// and only *waits* for the totals of row r, which every warp published one iteration earlier:
// warps may drift a row apart, so the lookups of one overlap the reductions of another.
//   iteration r of a warp:  lookups(r + 1) -> wait sum[r] -> [thread 0: refill the slot of row r]
//                           -> totals(r), coefficient -> dot(r + 1), butterfly, publish(r + 1)
//                           -> column sums += coefficient * values(r)
// sum[b], b = r mod 4: mbarrier with one arrival per warp; totals buffer sc[b][16].  A warp
// overwrites sc[(r + 1) mod 4] only after it saw sum[r] complete, i.e. after every warp has
// published row r, which each does after reading the totals of row r - 1 >= r - 3.  "Every warp
// has published row r" also means every warp has the values of row r in registers (the
// published number depends on all of them), so its ring slot can be refilled.
constexpr int kSumBufs = 4;

template <int NC, int THREADS = kPassThreads, bool kPairs = false>
__global__ void __launch_bounds__(THREADS, 1)
em_pass_coded_v3_kernel(const unsigned char *__restrict__ rows, uint32_t row_bytes, int64_t ld,
                        int64_t n_rows, const double *__restrict__ weights,
                        const double *__restrict__ pi0, const double *__restrict__ pi1,
                        EmState *__restrict__ st, double *__restrict__ partials, int n_stages,
                        int accumulate) {
    static_assert(kPassWarps == 16 && kSumBufs * kPassWarps <= 2 * kPassWarps * kPassGroup &&
                      THREADS % 32 == 0 && THREADS <= kPassThreads && (!kPairs || NC <= 8),
                  "totals buffers live in the scratch area of the two-row kernels");
    constexpr int kWarps = THREADS / 32;
    pdl_launch_dependents();  // the tail kernel may be scheduled as SMs drain

    MXB_DYN_SHARED __align__(128) unsigned char smem_raw[];
    double *scratch = reinterpret_cast<double *>(
        smem_raw + (((size_t)n_stages * row_bytes + 127) & ~(size_t)127));
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 2 * kPassWarps * kPassGroup);
    uint64_t *sum = full + 16;   // em_pack_rows: at most 16 stages, 256 spare bytes behind them

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t r_begin = n_rows * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = n_rows * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(r_end - r_begin);
    const unsigned char *my_rows = rows + (size_t)r_begin * row_bytes;
    const double *my_w = weights + r_begin;
    const uint32_t stages_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(full);
    const uint32_t sum_u32 = smem_u32(sum);

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        for (int b = 0; b < kSumBufs; ++b) mbar_init(&sum[b], kWarps);
        mbar_init_fence();
    }
    // totals slots of the warps a smaller CTA does not have (read by the 16-lane butterfly)
    if (THREADS < kPassThreads && tid < kSumBufs * kPassWarps && (tid & 15) >= kWarps)
        scratch[tid] = 0.0;
    __syncthreads();
    if (tid == 0) {
        for (int q = 0; q < n_my && q < n_stages; ++q) {
            mbar_expect_tx(&full[q], row_bytes);
            bulk_load(smem_raw + (size_t)q * row_bytes, my_rows + (size_t)q * row_bytes, row_bytes,
                      &full[q]);
        }
    }
    pdl_wait();  // proportions and control block of the previous iteration are final
    if (st->done) {
        for (int q = 0; q < n_my && q < n_stages; ++q) mbar_wait_u32(full_u32 + 8u * (uint32_t)q, 0u);
        return;
    }
    const double *__restrict__ pi = st->cur ? pi1 : pi0;

    const int n_chunks = (int)(ld >> 1);
    double2 pr[NC], tr[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * THREADS;
        pr[k] = (c < n_chunks) ? reinterpret_cast<const double2 *>(pi)[c] : make_double2(0.0, 0.0);
        tr[k] = make_double2(0.0, 0.0);
    }
    const bool last_live = tid + (NC - 1) * THREADS < n_chunks;  // only chunk NC-1 can be ragged

    int bad = 0;
    // the values of a row: table lookups of this thread's 2 NC cells in ring slot s
    auto lookups = [&](double2 (&lv)[NC], const int s) {
        if (kPairs) {   // chunk-coded record: 8 code bytes per thread, table of double2
            const uint32_t rec_u32 = stages_u32 + (uint32_t)s * row_bytes;
            const uint2 cw = lds_v2_u32(rec_u32 + (uint32_t)tid * 8u);
            const uint32_t tab_u32 = rec_u32 + (uint32_t)(THREADS * 8);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const unsigned word = (k < 4) ? cw.x : cw.y;
                const unsigned off = ((word >> (8 * (k & 3))) & 0xFFu) << 4;
                lv[k] = lds_v2_f64(tab_u32 + off);
            }
            return;
        }
        const unsigned char *srec = smem_raw + (size_t)s * row_bytes;
        const uint16_t *codes = reinterpret_cast<const uint16_t *>(srec) + tid;
        const double *tab = reinterpret_cast<const double *>(srec + ld);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (k < NC - 1 || last_live) {
                const unsigned cc = codes[k * THREADS];   // cells 2c, 2c + 1
                lv[k] = make_double2(tab[cc & 0xFFu], tab[cc >> 8]);
            } else {
                lv[k] = make_double2(0.0, 0.0);
            }
        }
    };
    // this warp's part of row r's dot product -> sc[r mod 4][warp], one arrival on sum[r mod 4]
    auto publish = [&](const double2 (&lv)[NC], const int r) {
        double dx = 0.0, dy = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            dx = fma(lv[k].x, pr[k].x, dx);
            dy = fma(lv[k].y, pr[k].y, dy);
        }
        double v = dx + dy;
        v += shfl_xor_f64(v, 16);
        v += shfl_xor_f64(v, 8);
        v += shfl_xor_f64(v, 4);
        v += shfl_xor_f64(v, 2);
        v += shfl_xor_f64(v, 1);
        if (lane == 0) {
            const int b = r & (kSumBufs - 1);
            scratch[b * kPassWarps + warp] = v;
            mbar_arrive_u32(sum_u32 + 8u * (uint32_t)b);   // release: the store above is visible
        }
    };

    if (n_my > 0) {
        double2 lv0[NC], lv1[NC];
        int s_next = 0;            // ring slot and parity of the row whose values are fetched next
        uint32_t ph_next = 0;
        mbar_wait_u32(full_u32, 0u);
        lookups(lv0, 0);
        if (++s_next == n_stages) { s_next = 0; ph_next ^= 1u; }
        publish(lv0, 0);
        // one row: `cur` holds the values of row r, `nxt` receives those of row r + 1
        auto row_step = [&](double2 (&cur)[NC], double2 (&nxt)[NC], const int r) {
            const double w_r = my_w[r];
            const bool more = r + 1 < n_my;
            const int s_cur = (s_next == 0 ? n_stages : s_next) - 1;   // slot of row r
            if (more) {
                mbar_wait_u32(full_u32 + 8u * (uint32_t)s_next, ph_next);
                lookups(nxt, s_next);
                if (++s_next == n_stages) { s_next = 0; ph_next ^= 1u; }
            }
            const int b = r & (kSumBufs - 1);
            mbar_wait_u32(sum_u32 + 8u * (uint32_t)b, (uint32_t)(r >> 2) & 1u);
            if (tid == 0) {
                const int q = r + n_stages;
                if (q < n_my) {
                    const uint32_t bar = full_u32 + 8u * (uint32_t)s_cur;
                    mbar_expect_tx_u32(bar, row_bytes);
                    bulk_load_u32(stages_u32 + (uint32_t)s_cur * row_bytes,
                                  my_rows + (size_t)q * row_bytes, row_bytes, bar);
                }
            }
            // 16 warp totals of row r: one per lane of each half-warp
            double t = scratch[b * kPassWarps + (lane & 15)];
            t += shfl_xor_f64(t, 8);
            t += shfl_xor_f64(t, 4);
            t += shfl_xor_f64(t, 2);
            t += shfl_xor_f64(t, 1);
            double coef = 0.0;
            if (w_r != 0.0) {
                coef = w_r / t;
                bad |= (t == 0.0);
            }
            if (more) publish(nxt, r + 1);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                tr[k].x = fma(coef, cur[k].x, tr[k].x);
                tr[k].y = fma(coef, cur[k].y, tr[k].y);
            }
        };
        int r = 0;
        for (; r + 1 < n_my; r += 2) {
            row_step(lv0, lv1, r);
            row_step(lv1, lv0, r + 1);
        }
        if (r < n_my) row_step(lv0, lv1, r);
    }

    double2 *out = reinterpret_cast<double2 *>(partials + (size_t)blockIdx.x * ld);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * THREADS;
        if (c < n_chunks) {
            if (accumulate) {
                const double2 prev = out[c];
                out[c] = make_double2(prev.x + tr[k].x, prev.y + tr[k].y);
            } else {
                out[c] = tr[k];
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0 && warp == 0) atomicAdd(&st->bad, 1);
}

// ---- fused E+M pass for two restarts at once ---------------------------------------
// Restarts of one run_em call share the matrix (em.py:117-156), so two of them can
// share each read of L: the pass is HBM-bound and the fp64 pipe is ~20 % busy.  Same
// streaming structure as em_pass_fast_kernel; a thread keeps the proportions and the
// column sums of BOTH restarts in registers (96 of its 128), so a staged row is read
// from shared memory twice -- once for the two dot products, once for the two
// column-sum updates -- and a second block barrier per row pair releases the stages.
// The four (row, restart) dot products of a row pair ride one butterfly.
template <int NC, bool kAccumulate = false>
__global__ void __launch_bounds__(kPassThreads, 1)
em_pass_pair_kernel(const double *__restrict__ lin, int64_t ld, int64_t n_rows,
                    const double *__restrict__ weights, const double *__restrict__ pi_a0,
                    const double *__restrict__ pi_a1, const double *__restrict__ pi_b0,
                    const double *__restrict__ pi_b1, EmState *__restrict__ st,
                    double *__restrict__ partials_a, double *__restrict__ partials_b,
                    int n_stages) {
    static_assert(kPassGroup == 2 && kPassWarps == 16, "reduction layout below");
    pdl_launch_dependents();

    MXB_DYN_SHARED __align__(128) unsigned char smem_raw[];
    const uint32_t row_bytes = (uint32_t)(ld * sizeof(double));
    double *scratch = reinterpret_cast<double *>(smem_raw + (size_t)n_stages * row_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 2 * kPassWarps * kPassGroup);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t r_begin = n_rows * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = n_rows * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(r_end - r_begin);
    const unsigned char *my_rows = reinterpret_cast<const unsigned char *>(lin + r_begin * ld);
    const double *my_w = weights + r_begin;
    const uint32_t stages_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(full);

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {
        for (int q = 0; q < n_my && q < n_stages; ++q) {
            mbar_expect_tx(&full[q], row_bytes);
            bulk_load(smem_raw + (size_t)q * row_bytes, my_rows + (size_t)q * row_bytes, row_bytes,
                      &full[q]);
        }
    }

    pdl_wait();
    const int done_a = st[0].done, done_b = st[1].done;
    if (done_a && done_b) {
        for (int q = 0; q < n_my && q < n_stages; ++q) mbar_wait_u32(full_u32 + 8u * (uint32_t)q, 0u);
        return;
    }
    const double *__restrict__ pia = st[0].cur ? pi_a1 : pi_a0;
    const double *__restrict__ pib = st[1].cur ? pi_b1 : pi_b0;

    const int n_chunks = (int)(ld >> 1);
    double2 pa[NC], pb[NC], ta[NC], tb[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * kPassThreads;
        const bool in = c < n_chunks;
        pa[k] = in ? reinterpret_cast<const double2 *>(pia)[c] : make_double2(0.0, 0.0);
        pb[k] = in ? reinterpret_cast<const double2 *>(pib)[c] : make_double2(0.0, 0.0);
        ta[k] = make_double2(0.0, 0.0);
        tb[k] = make_double2(0.0, 0.0);
    }
    const bool last_live = tid + (NC - 1) * kPassThreads < n_chunks;

    int stage = 0;
    uint32_t phase = 0;
    int bad = 0;
    // quarter q4 = lane >> 3 owns pair (row g = q4 >> 1, restart = q4 & 1) after the butterfly
    const int q4 = lane >> 3;
    const bool mine_done = (q4 & 1) ? done_b != 0 : done_a != 0;
    for (int q0 = 0; q0 < n_my; q0 += kPassGroup) {
        const int q_mine = q0 + (q4 >> 1);
        const double w_mine = (q_mine < n_my) ? my_w[q_mine] : 0.0;
        int s_of[kPassGroup];
        double d[4];  // [row][restart]
        int s = stage;
        uint32_t ph = phase;
#pragma unroll
        for (int g = 0; g < kPassGroup; ++g) {
            s_of[g] = s;
            double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
            if (q0 + g < n_my) {
                mbar_wait_u32(full_u32 + 8u * (uint32_t)s, ph);
                const double2 *srow = reinterpret_cast<const double2 *>(
                    smem_raw + (size_t)s * row_bytes) + tid;
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    if (k < NC - 1 || last_live) {
                        const double2 l = srow[k * kPassThreads];
                        ax = fma(l.x, pa[k].x, ax);
                        ay = fma(l.y, pa[k].y, ay);
                        bx = fma(l.x, pb[k].x, bx);
                        by = fma(l.y, pb[k].y, by);
                    }
                }
            }
            d[2 * g] = ax + ay;
            d[2 * g + 1] = bx + by;
            if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        // four sums in one butterfly: halves keep a row, quarters keep a restart
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        double e0 = (up16 ? d[2] : d[0]) + shfl_xor_f64(up16 ? d[0] : d[2], 16);
        double e1 = (up16 ? d[3] : d[1]) + shfl_xor_f64(up16 ? d[1] : d[3], 16);
        double v = (up8 ? e1 : e0) + shfl_xor_f64(up8 ? e0 : e1, 8);
        v += shfl_xor_f64(v, 4);
        v += shfl_xor_f64(v, 2);
        v += shfl_xor_f64(v, 1);
        // scratch[pair q4][warp]; the loop's second barrier separates consecutive groups
        if ((lane & 7) == 0) scratch[q4 * kPassWarps + warp] = v;
        __syncthreads();
        // 16 warp totals per pair: lane reads two of them, 3-step butterfly inside its quarter
        double t = scratch[q4 * kPassWarps + (lane & 7)] + scratch[q4 * kPassWarps + 8 + (lane & 7)];
        t += shfl_xor_f64(t, 4);
        t += shfl_xor_f64(t, 2);
        t += shfl_xor_f64(t, 1);
        double coef_mine = 0.0;
        if (w_mine != 0.0) {
            coef_mine = w_mine / t;
            bad |= (t == 0.0 && !mine_done);
        }
        const double c0a = __shfl_sync(0xffffffffu, coef_mine, 0);
        const double c0b = __shfl_sync(0xffffffffu, coef_mine, 8);
        const double c1a = __shfl_sync(0xffffffffu, coef_mine, 16);
        const double c1b = __shfl_sync(0xffffffffu, coef_mine, 24);
#pragma unroll
        for (int g = 0; g < kPassGroup; ++g) {
            if (q0 + g < n_my) {
                const double ca = g ? c1a : c0a, cb = g ? c1b : c0b;
                const double2 *srow = reinterpret_cast<const double2 *>(
                    smem_raw + (size_t)s_of[g] * row_bytes) + tid;
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    if (k < NC - 1 || last_live) {
                        const double2 l = srow[k * kPassThreads];
                        ta[k].x = fma(ca, l.x, ta[k].x);
                        ta[k].y = fma(ca, l.y, ta[k].y);
                        tb[k].x = fma(cb, l.x, tb[k].x);
                        tb[k].y = fma(cb, l.y, tb[k].y);
                    }
                }
            }
        }
        __syncthreads();  // both staged rows have been read twice: release them
        if (tid == 0) {
#pragma unroll
            for (int g = 0; g < kPassGroup; ++g) {
                const int q = q0 + g + n_stages;
                if (q < n_my) {
                    const uint32_t bar = full_u32 + 8u * (uint32_t)s_of[g];
                    mbar_expect_tx_u32(bar, row_bytes);
                    bulk_load_u32(stages_u32 + (uint32_t)s_of[g] * row_bytes,
                                  my_rows + (size_t)q * row_bytes, row_bytes, bar);
                }
            }
        }
        stage = s;
        phase = ph;
    }

    double2 *out_a = reinterpret_cast<double2 *>(partials_a + (size_t)blockIdx.x * ld);
    double2 *out_b = reinterpret_cast<double2 *>(partials_b + (size_t)blockIdx.x * ld);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * kPassThreads;
        if (c < n_chunks) {
            if (kAccumulate) {   // on top of what the launch before this one left (coded sessions)
                const double2 qa = out_a[c], qb = out_b[c];
                out_a[c] = make_double2(qa.x + ta[k].x, qa.y + ta[k].y);
                out_b[c] = make_double2(qb.x + tb[k].x, qb.y + tb[k].y);
            } else {
                out_a[c] = ta[k];
                out_b[c] = tb[k];
            }
        }
    }
    if (bad) atomicOr(&st[(q4 & 1)].bad, 1);
}

// Two restarts per read of the chunk-coded records (em_pack_pairs_kernel, 512-thread layout):
// em_pass_pair_kernel with the two values of a chunk looked up in the row's table, in both
// sweeps over a staged row.  Writes the column sums; the fp64 pair pass over the dense rows
// (em_pass_pair_kernel<NC, true>) adds its own afterwards.  Experimental, MXB_EM_CODED_PAIRS=1.
template <int NC>
__global__ void __launch_bounds__(kPassThreads, 1)
em_pass_pair_coded_kernel(const unsigned char *__restrict__ rec, int64_t ld, int64_t n_rows,
                    const double *__restrict__ weights, const double *__restrict__ pi_a0,
                    const double *__restrict__ pi_a1, const double *__restrict__ pi_b0,
                    const double *__restrict__ pi_b1, EmState *__restrict__ st,
                    double *__restrict__ partials_a, double *__restrict__ partials_b,
                    int n_stages) {
    static_assert(kPassGroup == 2 && kPassWarps == 16, "reduction layout below");
    pdl_launch_dependents();

    MXB_DYN_SHARED __align__(128) unsigned char smem_raw[];
    constexpr uint32_t row_bytes = (uint32_t)pair_rec_bytes(kPassThreads);
    double *scratch = reinterpret_cast<double *>(smem_raw + (size_t)n_stages * row_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 2 * kPassWarps * kPassGroup);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t r_begin = n_rows * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = n_rows * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(r_end - r_begin);
    const unsigned char *my_rows = rec + (size_t)r_begin * row_bytes;
    const double *my_w = weights + r_begin;
    const uint32_t stages_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(full);

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {
        for (int q = 0; q < n_my && q < n_stages; ++q) {
            mbar_expect_tx(&full[q], row_bytes);
            bulk_load(smem_raw + (size_t)q * row_bytes, my_rows + (size_t)q * row_bytes, row_bytes,
                      &full[q]);
        }
    }

    pdl_wait();
    const int done_a = st[0].done, done_b = st[1].done;
    if (done_a && done_b) {
        for (int q = 0; q < n_my && q < n_stages; ++q) mbar_wait_u32(full_u32 + 8u * (uint32_t)q, 0u);
        return;
    }
    const double *__restrict__ pia = st[0].cur ? pi_a1 : pi_a0;
    const double *__restrict__ pib = st[1].cur ? pi_b1 : pi_b0;

    const int n_chunks = (int)(ld >> 1);
    double2 pa[NC], pb[NC], ta[NC], tb[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * kPassThreads;
        const bool in = c < n_chunks;
        pa[k] = in ? reinterpret_cast<const double2 *>(pia)[c] : make_double2(0.0, 0.0);
        pb[k] = in ? reinterpret_cast<const double2 *>(pib)[c] : make_double2(0.0, 0.0);
        ta[k] = make_double2(0.0, 0.0);
        tb[k] = make_double2(0.0, 0.0);
    }

    int stage = 0;
    uint32_t phase = 0;
    int bad = 0;
    // quarter q4 = lane >> 3 owns pair (row g = q4 >> 1, restart = q4 & 1) after the butterfly
    const int q4 = lane >> 3;
    const bool mine_done = (q4 & 1) ? done_b != 0 : done_a != 0;
    for (int q0 = 0; q0 < n_my; q0 += kPassGroup) {
        const int q_mine = q0 + (q4 >> 1);
        const double w_mine = (q_mine < n_my) ? my_w[q_mine] : 0.0;
        int s_of[kPassGroup];
        double d[4];  // [row][restart]
        int s = stage;
        uint32_t ph = phase;
#pragma unroll
        for (int g = 0; g < kPassGroup; ++g) {
            s_of[g] = s;
            double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
            if (q0 + g < n_my) {
                mbar_wait_u32(full_u32 + 8u * (uint32_t)s, ph);
                const uint32_t rec_u32 = stages_u32 + (uint32_t)s * row_bytes;
                uint2 cw;   // this thread's chunk codes of the row
                cw = lds_v2_u32(rec_u32 + (uint32_t)tid * 8u);
                const uint32_t tab_u32 = rec_u32 + (uint32_t)(kPassThreads * 8);
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    // a chunk past the end of the row has code 0 and proportions 0
                    const unsigned word = (k < 4) ? cw.x : cw.y;
                    const unsigned off = ((word >> (8 * (k & 3))) & 0xFFu) << 4;
                    double2 l;
                    l = lds_v2_f64(tab_u32 + off);
                    ax = fma(l.x, pa[k].x, ax);
                    ay = fma(l.y, pa[k].y, ay);
                    bx = fma(l.x, pb[k].x, bx);
                    by = fma(l.y, pb[k].y, by);
                }
            }
            d[2 * g] = ax + ay;
            d[2 * g + 1] = bx + by;
            if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        // four sums in one butterfly: halves keep a row, quarters keep a restart
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        double e0 = (up16 ? d[2] : d[0]) + shfl_xor_f64(up16 ? d[0] : d[2], 16);
        double e1 = (up16 ? d[3] : d[1]) + shfl_xor_f64(up16 ? d[1] : d[3], 16);
        double v = (up8 ? e1 : e0) + shfl_xor_f64(up8 ? e0 : e1, 8);
        v += shfl_xor_f64(v, 4);
        v += shfl_xor_f64(v, 2);
        v += shfl_xor_f64(v, 1);
        // scratch[pair q4][warp]; the loop's second barrier separates consecutive groups
        if ((lane & 7) == 0) scratch[q4 * kPassWarps + warp] = v;
        __syncthreads();
        // 16 warp totals per pair: lane reads two of them, 3-step butterfly inside its quarter
        double t = scratch[q4 * kPassWarps + (lane & 7)] + scratch[q4 * kPassWarps + 8 + (lane & 7)];
        t += shfl_xor_f64(t, 4);
        t += shfl_xor_f64(t, 2);
        t += shfl_xor_f64(t, 1);
        double coef_mine = 0.0;
        if (w_mine != 0.0) {
            coef_mine = w_mine / t;
            bad |= (t == 0.0 && !mine_done);
        }
        const double c0a = __shfl_sync(0xffffffffu, coef_mine, 0);
        const double c0b = __shfl_sync(0xffffffffu, coef_mine, 8);
        const double c1a = __shfl_sync(0xffffffffu, coef_mine, 16);
        const double c1b = __shfl_sync(0xffffffffu, coef_mine, 24);
#pragma unroll
        for (int g = 0; g < kPassGroup; ++g) {
            if (q0 + g < n_my) {
                const double ca = g ? c1a : c0a, cb = g ? c1b : c0b;
                const uint32_t rec_u32 = stages_u32 + (uint32_t)s_of[g] * row_bytes;
                uint2 cw;
                cw = lds_v2_u32(rec_u32 + (uint32_t)tid * 8u);
                const uint32_t tab_u32 = rec_u32 + (uint32_t)(kPassThreads * 8);
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const unsigned word = (k < 4) ? cw.x : cw.y;
                    const unsigned off = ((word >> (8 * (k & 3))) & 0xFFu) << 4;
                    double2 l;
                    l = lds_v2_f64(tab_u32 + off);
                    ta[k].x = fma(ca, l.x, ta[k].x);
                    ta[k].y = fma(ca, l.y, ta[k].y);
                    tb[k].x = fma(cb, l.x, tb[k].x);
                    tb[k].y = fma(cb, l.y, tb[k].y);
                }
            }
        }
        __syncthreads();  // both staged rows have been read twice: release them
        if (tid == 0) {
#pragma unroll
            for (int g = 0; g < kPassGroup; ++g) {
                const int q = q0 + g + n_stages;
                if (q < n_my) {
                    const uint32_t bar = full_u32 + 8u * (uint32_t)s_of[g];
                    mbar_expect_tx_u32(bar, row_bytes);
                    bulk_load_u32(stages_u32 + (uint32_t)s_of[g] * row_bytes,
                                  my_rows + (size_t)q * row_bytes, row_bytes, bar);
                }
            }
        }
        stage = s;
        phase = ph;
    }

    double2 *out_a = reinterpret_cast<double2 *>(partials_a + (size_t)blockIdx.x * ld);
    double2 *out_b = reinterpret_cast<double2 *>(partials_b + (size_t)blockIdx.x * ld);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * kPassThreads;
        if (c < n_chunks) { out_a[c] = ta[k]; out_b[c] = tb[k]; }
    }
    if (bad) atomicOr(&st[(q4 & 1)].bad, 1);
}

// ---- general path: any shape, two passes over L -------------------------------
// coef_i = w_i / sum_j L_ij pi_j, one warp per row.
__global__ void __launch_bounds__(256)
em_rowdot_kernel(const double *__restrict__ lin, int64_t ld, int64_t n_rows,
                 const double *__restrict__ weights, const double *__restrict__ pi0,
                 const double *__restrict__ pi1, EmState *__restrict__ st,
                 double *__restrict__ coef) {
    if (st->done) return;
    const double2 *__restrict__ pi = reinterpret_cast<const double2 *>(st->cur ? pi1 : pi0);
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    const int64_t n_chunks = ld >> 1;
    int bad = 0;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const double2 *row = reinterpret_cast<const double2 *>(lin + r * ld);
        double dot = 0.0;
        for (int64_t c = lane; c < n_chunks; c += 32) {
            const double2 l = row[c];
            const double2 p = pi[c];
            dot = fma(l.x, p.x, dot);
            dot = fma(l.y, p.y, dot);
        }
        dot = warp_sum(dot);
        if (lane == 0) {
            const double w = weights[r];
            double c = 0.0;
            if (w != 0.0) { c = w / dot; bad |= (dot == 0.0); }
            coef[r] = c;
        }
    }
    if (bad) atomicAdd(&st->bad, 1);
}

// partial T_j over a row range; thread owns one double2 column chunk.
__global__ void __launch_bounds__(256)
em_colacc_kernel(const double *__restrict__ lin, int64_t ld, int64_t n_rows,
                 const double *__restrict__ coef, const EmState *__restrict__ st,
                 double *__restrict__ partials) {
    if (st->done) return;
    const int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x;
    if (c >= (ld >> 1)) return;
    const int64_t r_begin = n_rows * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = n_rows * (int64_t)(blockIdx.x + 1) / gridDim.x;
    double2 t = make_double2(0.0, 0.0);
    int64_t r = r_begin;
    for (; r + 4 <= r_end; r += 4) {
        double2 l[4];
        double cf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            l[u] = reinterpret_cast<const double2 *>(lin + (r + u) * ld)[c];
            cf[u] = coef[r + u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            t.x = fma(cf[u], l[u].x, t.x);
            t.y = fma(cf[u], l[u].y, t.y);
        }
    }
    for (; r < r_end; ++r) {
        const double2 l = reinterpret_cast<const double2 *>(lin + r * ld)[c];
        const double cf = coef[r];
        t.x = fma(cf, l.x, t.x);
        t.y = fma(cf, l.y, t.y);
    }
    reinterpret_cast<double2 *>(partials + (size_t)blockIdx.x * ld)[c] = t;
}

// T_j = sum over CTAs of partials, fixed order.
__global__ void __launch_bounds__(256)
em_colreduce_kernel(const double *__restrict__ partials, int n_part, int64_t ld,
                    const EmState *__restrict__ st, double *__restrict__ tsum) {
    if (st->done) return;
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= ld) return;
    double t = 0.0;
    for (int b = 0; b < n_part; ++b) t += partials[(size_t)b * ld + j];
    tsum[j] = t;
}

// M-step normalisation + convergence test (em.py:89 and :39-54), one CTA.
constexpr int kUpdThreads = 1024;
__global__ void __launch_bounds__(kUpdThreads)
em_update_kernel(const double *__restrict__ tsum, int64_t n_cols, int64_t ld,
                 double *__restrict__ lnp0, double *__restrict__ lnp1,
                 double *__restrict__ pi0, double *__restrict__ pi1,
                 EmState *__restrict__ st) {
    if (st->done) return;
    __shared__ double scratch[kUpdThreads / 32];
    const int cur = st->cur;
    const double *lnp_old = cur ? lnp1 : lnp0;
    const double *pi_old = cur ? pi1 : pi0;
    double *lnp_new = cur ? lnp0 : lnp1;
    double *pi_new = cur ? pi0 : pi1;

    double local = 0.0;
    for (int64_t j = threadIdx.x; j < n_cols; j += kUpdThreads) local += pi_old[j] * tsum[j];
    const double total = block_sum<kUpdThreads>(local, scratch);

    double dl = 0.0;
    for (int64_t j = threadIdx.x; j < n_cols; j += kUpdThreads) {
        const double p = pi_old[j];
        const double t = tsum[j];
        double ln_new;
        if (p >= 1e-290) ln_new = log(p * t / total);
        else ln_new = lnp_old[j] + log(t / total);  // pi underflowed: stay in log space
        const double p_new = exp(ln_new);
        lnp_new[j] = ln_new;
        pi_new[j] = p_new;
        dl += fabs(p_new - p);
    }
    const double delta = block_sum<kUpdThreads>(dl, scratch);
    if (threadIdx.x == 0) {
        st->delta = delta;
        const long long it = st->iters + 1;
        st->iters = it;
        if (delta < st->tol) st->done = 1;
        else if (it >= st->max_iter) st->done = 2;
        else st->cur = 1 - cur;
    }
}

#ifndef MXB_CPU_EMUL   // thread-block clusters and peer mailboxes: not modelled on the host
// ---- fused tail of an iteration (fast path) ------------------------------------
// One launch replaces em_colreduce_kernel + em_update_kernel: a single cluster of
// kFinCtas CTAs, one thread per column.  Each thread adds the per-CTA partial
// sums of its column in fixed order, the two scalars of the M-step (the
// normaliser sum_k pi_k T_k and the convergence distance sum_j |pi'_j - pi_j|,
// em.py:89 and :39-54) are reduced across the cluster through distributed shared
// memory, and CTA 0 advances the control block.  Deterministic: fixed summation
// orders, no atomics.
constexpr int kFinCtas = 8;
constexpr int kFinThreads = 1024;

// Sum of one double per thread over the whole cluster; every thread gets it.
// slots: kFinCtas doubles in *every* CTA's shared memory, wsum: kFinThreads/32.
__device__ __forceinline__ double cluster_sum(double v, double *wsum, double *slots) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = wsum[lane];  // kFinThreads / 32 == 32 warps
        t = warp_sum(t);
        if (lane < kFinCtas) {
            // lane r publishes this CTA's total in CTA r's slot array
            double *remote = cluster.map_shared_rank(slots, lane);
            remote[cluster.block_rank()] = t;
        }
    }
    cluster.sync();
    double total = 0.0;
#pragma unroll
    for (int r = 0; r < kFinCtas; ++r) total += slots[r];
    return total;
}

// Multi-GPU form of the tail (kP2P): the ranks' column sums are exchanged through
// peer memory instead of a separate collective.  Every rank stores its H sums
// into slot (seq & 1) of each peer's inbox (NVLink P2P stores), fences, and
// raises its flag there with the launch sequence number; it then waits until all
// `world` flags in its own block carry that number and adds the inbox rows in
// rank order, so that every rank forms bit-identical totals and takes the same
// convergence decision.  Two slots suffice: a rank can be at most one launch
// ahead of a peer, because it needs that peer's flag to finish a launch.
struct P2PArgs {
    int world, rank;
    unsigned char *block[kP2PMaxWorld];  // [r] = rank r's mailbox block (own block at [rank])
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kP2PTimeoutNs = 120ull * 1000000000ull;

template <bool kP2P>
__global__ void __cluster_dims__(kFinCtas, 1, 1) __launch_bounds__(kFinThreads)
em_finish_kernel(const double *partials, int n_part, int64_t n_cols, int64_t ld,
                 double *lnp0, double *lnp1, double *pi0, double *pi1, EmState *st, P2PArgs pa) {
    static_assert(kFinThreads == 1024, "cluster_sum assumes 32 warps");
    pdl_wait();               // the pass kernel's partial sums are complete and visible
    pdl_launch_dependents();  // the next pass may prime its ring while this tail runs
    {   // blockIdx.y = restart slot of a batched session (one cluster per slot)
        const size_t slot = blockIdx.y;
        st += slot;
        partials += slot * (size_t)n_part * ld;
        lnp0 += slot * ld; lnp1 += slot * ld;
        pi0 += slot * ld; pi1 += slot * ld;
    }
    if (st->done) return;  // same answer in every CTA: the block below is the only writer
    __shared__ double wsum[2][kFinThreads / 32];
    __shared__ double slots[2][kFinCtas];
    __shared__ int s_timeout;
    const int cur = st->cur;
    const double *lnp_old = cur ? lnp1 : lnp0;
    const double *pi_old = cur ? pi1 : pi0;
    double *lnp_new = cur ? lnp0 : lnp1;
    double *pi_new = cur ? pi0 : pi1;

    // Column sums.  A CTA owns ld/16 column pairs; its 1024 threads split the per-CTA
    // partials of a pair into n_grp contiguous ranges (three at H=5408) that are summed
    // concurrently and then added in range order: fixed order, a third of the latency.
    __shared__ double2 red[kFinThreads];
    const int ppc = (int)(ld >> 4);                       // column pairs per CTA
    const int n_grp = min(8, kFinThreads / ppc);
    const int pair_local = threadIdx.x % ppc, grp = threadIdx.x / ppc;
    const int64_t c = (int64_t)blockIdx.x * ppc + pair_local;   // columns 2c, 2c+1
    if (grp < n_grp) {
        const int b0 = (int)((int64_t)n_part * grp / n_grp), b1 = (int)((int64_t)n_part * (grp + 1) / n_grp);
        const double2 *col = reinterpret_cast<const double2 *>(partials) + c;
        const size_t stride = (size_t)(ld >> 1);
        double2 acc = make_double2(0.0, 0.0);
        int b = b0;
        for (; b + 4 <= b1; b += 4) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = col[(size_t)(b + u) * stride];
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
        }
        for (; b < b1; ++b) { const double2 v = col[(size_t)b * stride]; acc.x += v.x; acc.y += v.y; }
        red[grp * ppc + pair_local] = acc;
    }
    __syncthreads();
    const bool owner = grp == 0;                          // one thread per column pair from here on
    const bool live_x = owner && 2 * c < n_cols, live_y = owner && 2 * c + 1 < n_cols;
    double2 t = make_double2(0.0, 0.0);
    if (owner) {
        for (int g = 0; g < n_grp; ++g) { t.x += red[g * ppc + pair_local].x; t.y += red[g * ppc + pair_local].y; }
    }
    // a peer CTA's shared memory may only be written once that CTA is known to have
    // started: one cluster barrier before the first distributed-shared-memory store
    cooperative_groups::this_cluster().sync();
    unsigned long long seq = 0;
    if (kP2P) {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        const int W = pa.world;
        unsigned long long *my_flags = reinterpret_cast<unsigned long long *>(pa.block[pa.rank]);
        unsigned long long *my_seq = my_flags + 2 * kP2PMaxWorld;
        if (threadIdx.x == 0) s_timeout = 0;
        seq = *my_seq + 1;  // advanced by CTA 0 at the very end of this launch
        const int slot = (int)(seq & 1ull);
        if (owner) {
            for (int r = 0; r < W; ++r) {
                double2 *inbox = reinterpret_cast<double2 *>(pa.block[r] + kP2PInboxOffset);
                inbox[((size_t)slot * kP2PMaxWorld + pa.rank) * (kP2PMaxLd / 2) + c] = t;
            }
        }
        __threadfence_system();
        cluster.sync();  // every store of this rank is fenced
        if (blockIdx.x == 0 && threadIdx.x < W) {
            const int r = threadIdx.x;
            unsigned long long *peer_flags = reinterpret_cast<unsigned long long *>(pa.block[r]);
            st_release_sys(&peer_flags[slot * kP2PMaxWorld + pa.rank], seq);
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(&my_flags[slot * kP2PMaxWorld + r]) != seq) {
                if (global_timer_ns() - t0 > kP2PTimeoutNs) { s_timeout = 1; break; }
            }
        }
        cluster.sync();  // all ranks' sums have landed in this rank's inbox
        if (owner) {
            const double2 *inbox = reinterpret_cast<const double2 *>(pa.block[pa.rank] + kP2PInboxOffset);
            t = make_double2(0.0, 0.0);
            for (int r = 0; r < W; ++r) {
                const double2 v = __ldcg(&inbox[((size_t)slot * kP2PMaxWorld + r) * (kP2PMaxLd / 2) + c]);
                t.x += v.x;
                t.y += v.y;
            }
        }
    }
    double2 p = make_double2(0.0, 0.0);
    if (owner) p = reinterpret_cast<const double2 *>(pi_old)[c];
    if (!live_x) p.x = 0.0;
    if (!live_y) p.y = 0.0;
    const double total = cluster_sum((live_x ? p.x * t.x : 0.0) + (live_y ? p.y * t.y : 0.0),
                                     wsum[0], slots[0]);

    double dl = 0.0;
    if (owner) {
        double2 ln_new = make_double2(-INFINITY, -INFINITY), p_new = make_double2(0.0, 0.0);
        if (live_x) {
            if (p.x >= 1e-290) ln_new.x = log(p.x * t.x / total);
            else ln_new.x = lnp_old[2 * c] + log(t.x / total);  // pi underflowed: stay in log space
            p_new.x = exp(ln_new.x);
            dl += fabs(p_new.x - p.x);
        }
        if (live_y) {
            if (p.y >= 1e-290) ln_new.y = log(p.y * t.y / total);
            else ln_new.y = lnp_old[2 * c + 1] + log(t.y / total);
            p_new.y = exp(ln_new.y);
            dl += fabs(p_new.y - p.y);
        }
        reinterpret_cast<double2 *>(lnp_new)[c] = ln_new;   // padding columns: (-inf, 0) as set_props left them
        reinterpret_cast<double2 *>(pi_new)[c] = p_new;
    }
    const double delta = cluster_sum(dl, wsum[1], slots[1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->delta = delta;
        const long long it = st->iters + 1;
        st->iters = it;
        if (kP2P) {
            unsigned long long *my_flags = reinterpret_cast<unsigned long long *>(pa.block[pa.rank]);
            my_flags[2 * kP2PMaxWorld] = seq;
        }
        if (kP2P && s_timeout) st->done = 3;  // a peer never showed up
        else if (delta < st->tol) st->done = 1;
        else if (it >= st->max_iter) st->done = 2;
        else st->cur = 1 - cur;
    }
}

// ln pi -> (ln pi, pi) device buffers, padding zeroed.
#endif  // MXB_CPU_EMUL

__global__ void em_set_props_kernel(const double *__restrict__ src, int64_t n_cols, int64_t ld,
                                    double *__restrict__ lnp, double *__restrict__ pi,
                                    double *__restrict__ lnp_other, double *__restrict__ pi_other) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    const bool in = j < n_cols;
    const double v = in ? src[j] : -INFINITY;
    lnp[j] = v;
    pi[j] = in ? exp(v) : 0.0;
    lnp_other[j] = -INFINITY;
    pi_other[j] = 0.0;
}

// ---- read matrix: Z = (M + ln pi) - logsumexp_row(M + ln pi) -------------------
// (em.py:80-83).  mode 1 folds into dst with numpy.logaddexp (em.py:156).
__device__ __forceinline__ double np_logaddexp(double x, double y) {
    if (x == y) return x + 0.693147180559945309417232121458176568;
    const double tmp = x - y;
    if (tmp > 0) return x + log1p(exp(-tmp));
    if (tmp <= 0) return y + log1p(exp(tmp));
    return tmp;  // NaN
}

constexpr int kMixThreads = 256;
__global__ void __launch_bounds__(kMixThreads)
read_mix_kernel(const double *__restrict__ m, int64_t n_rows, int64_t n_cols,
                const double *__restrict__ lnp, double *__restrict__ dst, int mode,
                double sub_log) {
    __shared__ double scratch[kMixThreads / 32];
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const double *row = m + r * n_cols;
        double mx = -INFINITY;
        double nan_flag = 0.0;  // fmax drops NaN; numpy.max propagates it
        for (int64_t j = threadIdx.x; j < n_cols; j += kMixThreads) {
            const double z = row[j] + lnp[j];
            if (z != z) nan_flag = 1.0;
            mx = fmax(mx, z);
        }
        nan_flag = block_sum<kMixThreads>(nan_flag, scratch);
        mx = block_max<kMixThreads>(mx, scratch);
        double lse;
        if (nan_flag > 0.0) {
            lse = NAN;
        } else if (isinf(mx)) {
            lse = mx;  // all -inf -> log(0); any +inf -> +inf (scipy's out_inf branch)
        } else {
            // scipy _logsumexp: s over non-max terms, m = number of max terms
            double s = 0.0, cnt = 0.0;
            for (int64_t j = threadIdx.x; j < n_cols; j += kMixThreads) {
                const double z = row[j] + lnp[j];
                if (z == mx) cnt += 1.0;
                else s += exp(z - mx);
            }
            s = block_sum<kMixThreads>(s, scratch);
            cnt = block_sum<kMixThreads>(cnt, scratch);
            lse = log1p(s / cnt) + log(cnt) + mx;
        }
        double *out = dst + r * n_cols;
        for (int64_t j = threadIdx.x; j < n_cols; j += kMixThreads) {
            double z = (row[j] + lnp[j]) - lse;
            if (mode == 1) z = np_logaddexp(out[j], z);
            if (sub_log != 0.0) z -= sub_log;
            out[j] = z;
        }
    }
}

// Cross-rank fold helpers: m <- exp(m - mx) ; m <- mx + log(m) - sub_log.
__global__ void fold_exp_kernel(double *__restrict__ m, const double *__restrict__ mx, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double a = mx[i];
        m[i] = isinf(a) ? (a < 0 ? 0.0 : 1.0) : exp(m[i] - a);
    }
}
__global__ void fold_log_kernel(double *__restrict__ m, const double *__restrict__ mx, int64_t n,
                                double sub_log) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double a = mx[i];
        m[i] = (isinf(a) ? a : a + log(m[i])) - sub_log;
    }
}

}  // namespace mxb
