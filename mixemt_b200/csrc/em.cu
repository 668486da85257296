// Kernel 2: the EM mixture-proportion fit.
//
// Replaces em_step / converged / run_em (reference mixemt/em.py:39-165) and the
// scipy.special.logsumexp calls inside them (em.py:82, :87-89).
//
// Formulation.  The reference works on Z_ij = M_ij + ln pi_j - lse_i in log
// space and makes ~30 passes over the N x H matrix per iteration (two
// max-shifted logsumexp reductions).  With m_i = max_j M_ij and
// L_ij = exp(M_ij - m_i) (computed once per run_em call, shared by all
// restarts), one iteration is
//
//     s_i   = sum_j L_ij pi_j                      (row dot product)
//     T_j   = sum_i (w_i / s_i) L_ij               (weighted column sum)
//     pi'_j = pi_j T_j / sum_k pi_k T_k            (M-step; == em.py:87-89)
//
// i.e. exactly two fp64 FMAs per cell and ONE read of L per iteration: the pass
// is bound by HBM bandwidth (8 B/cell), not by exp throughput.  ln pi is
// carried in log space next to pi (ln pi'_j = ln pi_j + log(T_j / total) once
// pi_j leaves the normal range), so dying components keep following the
// reference's trajectory after exp() underflows.  The read matrix the reference
// returns (em.py:130-136: responsibilities at the *previous* proportions) is
// produced in log space straight from M by read_mix_kernel, with scipy's
// logsumexp formula (log1p(s/m) + log(m) + max).
//
// The fused pass (em_pass_fast_kernel) is a persistent kernel, one CTA per SM,
// that streams its contiguous row range through a shared-memory ring filled by
// 1-D bulk async copies (TMA engine, cp.async.bulk + mbarrier complete_tx).
// Thread t owns the same 2*NC columns of every row: it reads each staged cell
// once into registers, contributes to the row's dot product, and after a
// block-wide reduction adds (w_i/s_i) * L_ij into its private column sums.
// Per-CTA column sums are written once at the end and reduced in a fixed order
// (deterministic: no floating-point atomics anywhere).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <vector>

#include <cooperative_groups.h>

#include <type_traits>

#include "common.cuh"

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
    extern __shared__ unsigned short cell_slot[];   // [ld] hash slot of every cell
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
    extern __shared__ unsigned short chunk_slot[];   // [ld / 2] hash slot of every chunk
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

    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                         : "=r"(cw.x), "=r"(cw.y) : "r"(rec_u32 + (uint32_t)tid * 8u) : "memory");
            const uint32_t tab_u32 = rec_u32 + (uint32_t)(THREADS * 8);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                // a chunk past the end of the row has code 0 and proportion 0: whatever the
                // table holds there adds nothing to the dot product, and its column sum is
                // never written
                const unsigned word = (k < 4) ? cw.x : cw.y;
                const unsigned off = ((word >> (8 * (k & 3))) & 0xFFu) << 4;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                             : "=d"(lv[g][k].x), "=d"(lv[g][k].y) : "r"(tab_u32 + off) : "memory");
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

__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int NC>
__global__ void __launch_bounds__(kPassThreads, 1)
em_pass_coded_v3_kernel(const unsigned char *__restrict__ rows, uint32_t row_bytes, int64_t ld,
                        int64_t n_rows, const double *__restrict__ weights,
                        const double *__restrict__ pi0, const double *__restrict__ pi1,
                        EmState *__restrict__ st, double *__restrict__ partials, int n_stages,
                        int accumulate) {
    static_assert(kPassWarps == 16 && kSumBufs * kPassWarps <= 2 * kPassWarps * kPassGroup,
                  "totals buffers live in the scratch area of the two-row kernels");
    pdl_launch_dependents();  // the tail kernel may be scheduled as SMs drain

    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        for (int b = 0; b < kSumBufs; ++b) mbar_init(&sum[b], kPassWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
        for (int q = 0; q < n_my && q < n_stages; ++q) mbar_wait_u32(full_u32 + 8u * (uint32_t)q, 0u);
        return;
    }
    const double *__restrict__ pi = st->cur ? pi1 : pi0;

    const int n_chunks = (int)(ld >> 1);
    double2 pr[NC], tr[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = tid + k * kPassThreads;
        pr[k] = (c < n_chunks) ? reinterpret_cast<const double2 *>(pi)[c] : make_double2(0.0, 0.0);
        tr[k] = make_double2(0.0, 0.0);
    }
    const bool last_live = tid + (NC - 1) * kPassThreads < n_chunks;  // only chunk NC-1 can be ragged

    int bad = 0;
    // the values of a row: table lookups of this thread's 2 NC cells in ring slot s
    auto lookups = [&](double2 (&lv)[NC], const int s) {
        const unsigned char *srec = smem_raw + (size_t)s * row_bytes;
        const uint16_t *codes = reinterpret_cast<const uint16_t *>(srec) + tid;
        const double *tab = reinterpret_cast<const double *>(srec + ld);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (k < NC - 1 || last_live) {
                const unsigned cc = codes[k * kPassThreads];   // cells 2c, 2c + 1
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

    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                             : "=r"(cw.x), "=r"(cw.y) : "r"(rec_u32 + (uint32_t)tid * 8u) : "memory");
                const uint32_t tab_u32 = rec_u32 + (uint32_t)(kPassThreads * 8);
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    // a chunk past the end of the row has code 0 and proportions 0
                    const unsigned word = (k < 4) ? cw.x : cw.y;
                    const unsigned off = ((word >> (8 * (k & 3))) & 0xFFu) << 4;
                    double2 l;
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                                 : "=d"(l.x), "=d"(l.y) : "r"(tab_u32 + off) : "memory");
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
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                             : "=r"(cw.x), "=r"(cw.y) : "r"(rec_u32 + (uint32_t)tid * 8u) : "memory");
                const uint32_t tab_u32 = rec_u32 + (uint32_t)(kPassThreads * 8);
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const unsigned word = (k < 4) ? cw.x : cw.y;
                    const unsigned off = ((word >> (8 * (k & 3))) & 0xFFu) << 4;
                    double2 l;
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                                 : "=d"(l.x), "=d"(l.y) : "r"(tab_u32 + off) : "memory");
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

using namespace mxb;

struct mxb_em {
    mxb_ctx *ctx = nullptr;
    const mxb_matrix *mat = nullptr;
    int64_t n_rows = 0, n_cols = 0, ld = 0;
    bool sharded = false;
    double *lin = nullptr;        // [n_rows][ld]
    double *weights = nullptr;    // [n_rows]
    double *coef = nullptr;       // [n_rows] (general path)
    double *lnp[2] = {nullptr, nullptr};
    double *pi[2] = {nullptr, nullptr};
    double *partials = nullptr;   // [n_part][ld]
    double *tsum = nullptr;       // [ld]
    double *props_in = nullptr;   // [n_cols] staging for set_lnprops
    EmState *state = nullptr;     // device
    EmState *host_state = nullptr;  // pinned, 2 slots
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    int n_part = 0;
    // fast path
    bool fast = false;
    int nc = 0;
    int n_stages = 0;
    size_t smem_bytes = 0;
    int grid_fast = 0;
    // general path
    int row_blocks = 0, col_blocks = 0;
    bool fused_tail = false;  // colreduce + update in one cluster launch (em_finish_kernel)
    // Restart slots: a batched session (run_em with n_multi > 1) iterates two restarts per
    // read of L.  Slot s lives at lnp[b] + s*ld, pi[b] + s*ld, partials + s*n_part*ld, state + s.
    int n_slots = 1;
    // Dictionary-coded rows (em_pack_kernel): when `coded` the pass reads `rec` (all rows,
    // dense rows as empty records) and `dense_lin` (the gathered dense rows) instead of `lin`.
    bool coded = false;
    int pair_threads = 0;             // != 0: records hold chunk dictionaries laid out for a pass
                                      // kernel of that many threads (em_pack_pairs_kernel)
    unsigned char *rec = nullptr;     // [n_coded][rec_bytes]
    int64_t n_coded = 0;              // rows of `rec`: all rows, or the coded ones only (compact)
    size_t rec_bytes = 0;
    double *w_coded = nullptr;        // [n_rows] weight, 0 for dense rows
    double *dense_lin = nullptr;      // [n_dense][ld]
    double *w_dense = nullptr;        // [n_dense]
    int64_t n_dense = 0;
    int coded_stages = 0;
    size_t coded_smem = 0;
    int grid_dense = 0;
    unsigned char *small = nullptr;  // one device block behind weights ... state (fewer driver calls)
    bool zero_iter = false;  // last iterate() ran no iteration
};

namespace mxb {

typedef void (*pass_fn)(const unsigned char *, uint32_t, int64_t, int64_t, const double *,
                        const double *, const double *, EmState *, double *, int, int);

static pass_fn pick_pass(int nc) {
    switch (nc) {
        case 1: return em_pass_fast_kernel<1>;
        case 2: return em_pass_fast_kernel<2>;
        case 3: return em_pass_fast_kernel<3>;
        case 4: return em_pass_fast_kernel<4>;
        case 5: return em_pass_fast_kernel<5>;
        case 6: return em_pass_fast_kernel<6>;
        case 7: return em_pass_fast_kernel<7>;
        case 8: return em_pass_fast_kernel<8>;
    }
    return nullptr;
}
// The coded pass and its CTA size.  Default: em_pass_coded_kernel, 512 threads, chunk count
// `nc`.  Experimental, not yet run on a GPU:
// MXB_EM_CODED_V3=1 (pipelined rows, no block barrier) and MXB_EM_CODED_T384=1 (the same
// kernel with 384 threads: 16 instead of 12 cells per thread at H = 5408, so the per-row-pair
// reduction, division and ring bookkeeping of a warp are spread over a third more cells).
struct CodedPass {
    pass_fn fn;
    int threads;
};
static CodedPass pick_pass_coded(int nc, int64_t ld, int pair_threads) {
    if (pair_threads == 384) {
        switch ((int)ceil_div(ld / 2, 384)) {
            case 1: return {em_pass_coded_pairs_kernel<1, 384>, 384};
            case 2: return {em_pass_coded_pairs_kernel<2, 384>, 384};
            case 3: return {em_pass_coded_pairs_kernel<3, 384>, 384};
            case 4: return {em_pass_coded_pairs_kernel<4, 384>, 384};
            case 5: return {em_pass_coded_pairs_kernel<5, 384>, 384};
            case 6: return {em_pass_coded_pairs_kernel<6, 384>, 384};
            case 7: return {em_pass_coded_pairs_kernel<7, 384>, 384};
            case 8: return {em_pass_coded_pairs_kernel<8, 384>, 384};
        }
        return {nullptr, 0};
    }
    if (pair_threads != 0) {
        switch (nc) {
            case 1: return {em_pass_coded_pairs_kernel<1>, kPassThreads};
            case 2: return {em_pass_coded_pairs_kernel<2>, kPassThreads};
            case 3: return {em_pass_coded_pairs_kernel<3>, kPassThreads};
            case 4: return {em_pass_coded_pairs_kernel<4>, kPassThreads};
            case 5: return {em_pass_coded_pairs_kernel<5>, kPassThreads};
            case 6: return {em_pass_coded_pairs_kernel<6>, kPassThreads};
            case 7: return {em_pass_coded_pairs_kernel<7>, kPassThreads};
            case 8: return {em_pass_coded_pairs_kernel<8>, kPassThreads};
        }
        return {nullptr, 0};
    }
    static const bool v3 = getenv("MXB_EM_CODED_V3") != nullptr;
    static const bool t384 = getenv("MXB_EM_CODED_T384") != nullptr;
    if (v3) {
        switch (nc) {
            case 1: return {em_pass_coded_v3_kernel<1>, kPassThreads};
            case 2: return {em_pass_coded_v3_kernel<2>, kPassThreads};
            case 3: return {em_pass_coded_v3_kernel<3>, kPassThreads};
            case 4: return {em_pass_coded_v3_kernel<4>, kPassThreads};
            case 5: return {em_pass_coded_v3_kernel<5>, kPassThreads};
            case 6: return {em_pass_coded_v3_kernel<6>, kPassThreads};
            case 7: return {em_pass_coded_v3_kernel<7>, kPassThreads};
            case 8: return {em_pass_coded_v3_kernel<8>, kPassThreads};
        }
        return {nullptr, 0};
    }
    if (t384) {
        switch ((int)ceil_div(ld / 2, 384)) {
            case 6: return {em_pass_coded_kernel<6, 384>, 384};
            case 7: return {em_pass_coded_kernel<7, 384>, 384};
            case 8: return {em_pass_coded_kernel<8, 384>, 384};
        }
        // other widths keep the 512-thread kernel
    }
    switch (nc) {
        case 1: return {em_pass_coded_kernel<1>, kPassThreads};
        case 2: return {em_pass_coded_kernel<2>, kPassThreads};
        case 3: return {em_pass_coded_kernel<3>, kPassThreads};
        case 4: return {em_pass_coded_kernel<4>, kPassThreads};
        case 5: return {em_pass_coded_kernel<5>, kPassThreads};
        case 6: return {em_pass_coded_kernel<6>, kPassThreads};
        case 7: return {em_pass_coded_kernel<7>, kPassThreads};
        case 8: return {em_pass_coded_kernel<8>, kPassThreads};
    }
    return {nullptr, 0};
}

typedef void (*pair_fn)(const double *, int64_t, int64_t, const double *, const double *,
                        const double *, const double *, const double *, EmState *, double *,
                        double *, int);
static pair_fn pick_pair(int nc) {
    switch (nc) {
        case 1: return em_pass_pair_kernel<1>;
        case 2: return em_pass_pair_kernel<2>;
        case 3: return em_pass_pair_kernel<3>;
        case 4: return em_pass_pair_kernel<4>;
        case 5: return em_pass_pair_kernel<5>;
        case 6: return em_pass_pair_kernel<6>;
    }
    return nullptr;
}
// the fp64 pair pass adding its column sums to what the coded pair pass left (dense rows)
static pair_fn pick_pair_accumulate(int nc) {
    switch (nc) {
        case 1: return em_pass_pair_kernel<1, true>;
        case 2: return em_pass_pair_kernel<2, true>;
        case 3: return em_pass_pair_kernel<3, true>;
        case 4: return em_pass_pair_kernel<4, true>;
        case 5: return em_pass_pair_kernel<5, true>;
        case 6: return em_pass_pair_kernel<6, true>;
    }
    return nullptr;
}
typedef void (*pair_coded_fn)(const unsigned char *, int64_t, int64_t, const double *,
                              const double *, const double *, const double *, const double *,
                              EmState *, double *, double *, int);
static pair_coded_fn pick_pair_coded(int nc) {
    switch (nc) {
        case 1: return em_pass_pair_coded_kernel<1>;
        case 2: return em_pass_pair_coded_kernel<2>;
        case 3: return em_pass_pair_coded_kernel<3>;
        case 4: return em_pass_pair_coded_kernel<4>;
        case 5: return em_pass_pair_coded_kernel<5>;
        case 6: return em_pass_pair_coded_kernel<6>;
    }
    return nullptr;
}
constexpr int kMaxPairNC = 6;   // 2 restarts x (pi + T) x NC double2 must fit 128 registers
constexpr int kMaxSlots = 2;

// Launch with the programmatic-stream-serialization attribute (see pdl_wait): the kernel may
// be scheduled before its predecessor in the stream has finished.  MXB_EM_NO_PDL=1 turns the
// attribute off (plain stream order).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t s, Args... args) {
    static const bool use_pdl = getenv("MXB_EM_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// One EM iteration on em->ctx->stream (no host sync).
static int enqueue_iteration(mxb_em *em, cudaEvent_t pass_begin = nullptr,
                             cudaEvent_t pass_end = nullptr) {
    mxb_ctx *ctx = em->ctx;
    cudaStream_t s = ctx->stream;
    if (pass_begin) MXB_CUDA(cudaEventRecord(pass_begin, s));
    if (em->n_slots == 2) {
        const size_t ps = (size_t)em->n_part * em->ld;
        if (em->coded) {
            // chunk-coded records of all rows, then the fp64 rows of the dense ones on top
            MXB_CUDA(launch_pdl(pick_pair_coded(em->nc), dim3(em->grid_fast), dim3(kPassThreads),
                                em->coded_smem, s, (const unsigned char *)em->rec, em->ld,
                                em->n_coded, em->w_coded, em->pi[0], em->pi[1], em->pi[0] + em->ld,
                                em->pi[1] + em->ld, em->state, em->partials, em->partials + ps,
                                em->coded_stages));
            ctx->launches += 1;
            if (em->n_dense > 0) {
                MXB_CUDA(launch_pdl(pick_pair_accumulate(em->nc), dim3(em->grid_fast),
                                    dim3(kPassThreads), em->smem_bytes, s, em->dense_lin, em->ld,
                                    em->n_dense, em->w_dense, em->pi[0], em->pi[1],
                                    em->pi[0] + em->ld, em->pi[1] + em->ld, em->state,
                                    em->partials, em->partials + ps, em->n_stages));
                ctx->launches += 1;
            }
        } else {
            MXB_CUDA(launch_pdl(pick_pair(em->nc), dim3(em->grid_fast), dim3(kPassThreads),
                                em->smem_bytes, s, em->lin, em->ld, em->n_rows, em->weights,
                                em->pi[0], em->pi[1], em->pi[0] + em->ld, em->pi[1] + em->ld,
                                em->state, em->partials, em->partials + ps, em->n_stages));
            ctx->launches += 1;
        }
        if (pass_end) MXB_CUDA(cudaEventRecord(pass_end, s));
        P2PArgs pa;
        memset(&pa, 0, sizeof(pa));
        MXB_CUDA(launch_pdl(em_finish_kernel<false>, dim3(kFinCtas, 2), dim3(kFinThreads), 0, s,
                            em->partials, em->n_part, em->n_cols, em->ld, em->lnp[0], em->lnp[1],
                            em->pi[0], em->pi[1], em->state, pa));
        ctx->launches += 1;
        return MXB_OK;
    }
    if (em->coded) {
        // dense rows first (their own small matrix), then the coded records of all rows on top;
        // both launches use the same grid, so partials[cta] is written, then added to
        if (em->n_dense > 0) {
            MXB_CUDA(launch_pdl(pick_pass(em->nc), dim3(em->grid_fast), dim3(kPassThreads),
                                em->smem_bytes, s, (const unsigned char *)em->dense_lin,
                                (uint32_t)(em->ld * sizeof(double)), em->ld, em->n_dense,
                                em->w_dense, em->pi[0], em->pi[1], em->state, em->partials,
                                em->n_stages, 0));
            ctx->launches += 1;
        }
        const CodedPass cp = pick_pass_coded(em->nc, em->ld, em->pair_threads);
        MXB_CUDA(launch_pdl(cp.fn, dim3(em->grid_fast), dim3(cp.threads),
                            em->coded_smem, s, (const unsigned char *)em->rec,
                            (uint32_t)em->rec_bytes, em->ld, em->n_coded, em->w_coded, em->pi[0],
                            em->pi[1], em->state, em->partials, em->coded_stages,
                            em->n_dense > 0 ? 1 : 0));
        ctx->launches += 1;
    } else if (em->fast) {
        MXB_CUDA(launch_pdl(pick_pass(em->nc), dim3(em->grid_fast), dim3(kPassThreads),
                            em->smem_bytes, s, (const unsigned char *)em->lin,
                            (uint32_t)(em->ld * sizeof(double)), em->ld, em->n_rows, em->weights,
                            em->pi[0], em->pi[1], em->state, em->partials, em->n_stages, 0));
        ctx->launches += 1;
    } else {
        const int rd_blocks = (int)std::max<int64_t>(
            1, std::min<int64_t>(ceil_div(em->n_rows, 8), (int64_t)ctx->num_sms * 8));
        em_rowdot_kernel<<<rd_blocks, 256, 0, s>>>(em->lin, em->ld, em->n_rows, em->weights,
                                                   em->pi[0], em->pi[1], em->state, em->coef);
        dim3 grid(em->row_blocks, em->col_blocks);
        em_colacc_kernel<<<grid, 256, 0, s>>>(em->lin, em->ld, em->n_rows, em->coef, em->state,
                                              em->partials);
        ctx->launches += 2;
    }
    if (pass_end) MXB_CUDA(cudaEventRecord(pass_end, s));
    if (em->fused_tail) {
        P2PArgs pa;
        memset(&pa, 0, sizeof(pa));
        if (em->sharded && ctx->world > 1) {
            pa.world = ctx->world;
            pa.rank = ctx->rank;
            for (int r = 0; r < ctx->world; ++r) pa.block[r] = ctx->p2p_block[r];
            MXB_CUDA(launch_pdl(em_finish_kernel<true>, dim3(kFinCtas), dim3(kFinThreads), 0, s,
                                em->partials, em->n_part, em->n_cols, em->ld, em->lnp[0],
                                em->lnp[1], em->pi[0], em->pi[1], em->state, pa));
        } else {
            MXB_CUDA(launch_pdl(em_finish_kernel<false>, dim3(kFinCtas), dim3(kFinThreads), 0, s,
                                em->partials, em->n_part, em->n_cols, em->ld, em->lnp[0],
                                em->lnp[1], em->pi[0], em->pi[1], em->state, pa));
        }
        ctx->launches += 1;
        return MXB_OK;
    }
    em_colreduce_kernel<<<(int)ceil_div(em->ld, 256), 256, 0, s>>>(em->partials, em->n_part,
                                                                  em->ld, em->state, em->tsum);
    ctx->launches += 1;
    if (em->sharded && ctx->world > 1) MXB_TRY(nccl_allreduce_sum_f64(ctx, em->tsum, em->ld));
    em_update_kernel<<<1, kUpdThreads, 0, s>>>(em->tsum, em->n_cols, em->ld, em->lnp[0],
                                               em->lnp[1], em->pi[0], em->pi[1], em->state);
    ctx->launches += 1;
    MXB_CUDA(cudaGetLastError());
    return MXB_OK;
}

// MXB_TIMING=1: wall-clock stage times of the one-call entry points on stderr.
struct StageTimer {
    bool on;
    cudaStream_t stream;
    std::chrono::steady_clock::time_point t;
    explicit StageTimer(cudaStream_t s) : on(getenv("MXB_TIMING") != nullptr), stream(s) {
        if (on) t = std::chrono::steady_clock::now();
    }
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(stream);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[mxb timing] %-28s %9.3f ms\n", what,
                std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

static int reset_state(mxb_em *em, long long max_iter, double tol, int slot = 0, int done = 0) {
    EmState st;
    memset(&st, 0, sizeof(st));
    st.max_iter = max_iter;
    st.tol = tol;
    st.done = done;
    // synchronous w.r.t. the host buffer: pageable copy of a stack struct
    MXB_CUDA(cudaMemcpyAsync(em->state + slot, &st, sizeof(st), cudaMemcpyHostToDevice,
                             em->ctx->stream));
    MXB_CUDA(cudaStreamSynchronize(em->ctx->stream));
    return MXB_OK;
}

}  // namespace mxb

extern "C" {

int mxb_em_destroy(mxb_em *em) {
    if (!em) return MXB_OK;
    cudaSetDevice(em->ctx->device);
    cudaStreamSynchronize(em->ctx->stream);
    dev_free(em->ctx, em->lin);
    dev_free(em->ctx, em->rec);
    dev_free(em->ctx, em->dense_lin);
    dev_free(em->ctx, em->small);
    for (int i = 0; i < 2; ++i)
        if (em->poll_ev[i]) cudaEventDestroy(em->poll_ev[i]);
    delete em;  // host_state points into the context's pinned scratch
    return MXB_OK;
}

}  // extern "C"

namespace mxb {

__global__ void em_reset_state_kernel(EmState *st, long long max_iter, double tol) {
    st->done = 0;
    st->cur = 0;
    st->bad = 0;
    st->iters = 0;
    st->max_iter = max_iter;
    st->tol = tol;
    st->delta = 0.0;
}

// Dictionary-code the rows of em->lin (see em_pack_kernel).  On success with enough
// codable rows the session switches to the coded pass and gives em->lin back; otherwise it
// stays as it is.  MXB_EM_NO_PACK=1 keeps the fp64 rows (cross-check).
static int em_pack_rows(mxb_em *em) {
    mxb_ctx *ctx = em->ctx;
    if (!em->fast || em->n_rows == 0 || getenv("MXB_EM_NO_PACK")) return MXB_OK;
    // restart pairs read fp64 rows unless the chunk dictionary is asked for (experimental)
    const bool two_slots = em->n_slots == 2;
    if (two_slots && getenv("MXB_EM_CODED_PAIRS") == nullptr) return MXB_OK;
    if (em->n_slots > 2) return MXB_OK;
    const size_t row_bytes = (size_t)em->ld * sizeof(double);
    // MXB_EM_CODED_PAIRS=1 (experimental): dictionaries of cell pairs, see em_pack_pairs_kernel
    // (with MXB_EM_CODED_T384=1 laid out for the 384-thread pass kernel where the row fits it)
    int pair_threads = 0;
    if (getenv("MXB_EM_CODED_PAIRS") != nullptr && em->nc <= 8)
        pair_threads = (!two_slots && getenv("MXB_EM_CODED_T384") != nullptr &&
                        ceil_div(em->ld / 2, 384) <= 8) ? 384 : kPassThreads;
    if (two_slots && pair_threads == 0) return MXB_OK;
    const bool pairs = pair_threads != 0;
    const size_t rec_bytes = pairs ? (size_t)pair_rec_bytes(pair_threads)
                                   : (size_t)em->ld + kDictSize * sizeof(double);
    constexpr int kMaxCodedStages = 16;
    const size_t fixed = 2 * kPassWarps * kPassGroup * sizeof(double) +
                         kMaxCodedStages * sizeof(uint64_t) + 256;
    if (ctx->smem_optin <= fixed) return MXB_OK;
    const int stages = (int)std::min<size_t>(kMaxCodedStages, (ctx->smem_optin - fixed) / rec_bytes);
    if (stages < kPassGroup + 1) return MXB_OK;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t n = (size_t)em->n_rows;
    unsigned char *rec = nullptr, *tmp = nullptr;
    cudaError_t e = dev_alloc(ctx, (void **)&rec, up(n * rec_bytes) + up(n * sizeof(double)));
    if (e == cudaSuccess)
        e = dev_alloc(ctx, (void **)&tmp, up(n * sizeof(int)) + up(n * sizeof(int64_t)) + 256);
    if (e != cudaSuccess) {   // not enough memory for the coded copy: keep the fp64 rows
        cudaGetLastError();
        dev_free(ctx, rec);
        dev_free(ctx, tmp);
        return MXB_OK;
    }
    double *w_coded = reinterpret_cast<double *>(rec + up(n * rec_bytes));
    int *flag = reinterpret_cast<int *>(tmp);
    int64_t *list = reinterpret_cast<int64_t *>(tmp + up(n * sizeof(int)));
    int64_t *d_count = reinterpret_cast<int64_t *>(tmp + up(n * sizeof(int)) + up(n * sizeof(int64_t)));
    int64_t n_dense = 0;
    double *dense = nullptr;
    const int grid = (int)std::min<int64_t>(em->n_rows, (int64_t)ctx->num_sms * 8);
    if (pairs)
        em_pack_pairs_kernel<<<grid, kPackThreads, (size_t)(em->ld / 2) * sizeof(unsigned short),
                               ctx->stream>>>(em->lin, em->n_rows, em->ld, em->weights, rec,
                                              pair_threads, flag, w_coded);
    else
        em_pack_kernel<<<grid, kPackThreads, (size_t)em->ld * sizeof(unsigned short), ctx->stream>>>(
            em->lin, em->n_rows, em->ld, em->weights, rec, (int64_t)rec_bytes, flag, w_coded);
    em_dense_list_kernel<<<1, 1024, 0, ctx->stream>>>(flag, em->n_rows, list, d_count);
    ctx->launches += 2;
    e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&n_dense, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    // worth it when the coded pass reads less than two thirds of the fp64 rows
    const bool worth = e == cudaSuccess &&
                       (double)n * rec_bytes + (double)n_dense * row_bytes < 0.66 * (double)n * row_bytes;
    if (worth && n_dense > 0) {
        e = dev_alloc(ctx, (void **)&dense, up((size_t)n_dense * row_bytes) + up((size_t)n_dense * sizeof(double)));
        if (e == cudaSuccess) {
            double *w_dense = reinterpret_cast<double *>((unsigned char *)dense + up((size_t)n_dense * row_bytes));
            const int g = (int)std::min<int64_t>(n_dense, (int64_t)ctx->num_sms * 8);
            em_gather_rows_kernel<<<g, 256, 0, ctx->stream>>>(em->lin, em->ld, em->weights, list,
                                                               n_dense, dense, w_dense);
            ctx->launches += 1;
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            em->w_dense = w_dense;
        }
    }
    int64_t n_coded = em->n_rows;
    if (e == cudaSuccess && worth && n_dense > 0 && getenv("MXB_EM_CODED_COMPACT") != nullptr) {
        // records of the coded rows only (the flags and the list buffer are reused)
        const int64_t n_keep = em->n_rows - n_dense;
        unsigned char *rec2 = nullptr;
        e = dev_alloc(ctx, (void **)&rec2, up((size_t)n_keep * rec_bytes) + up((size_t)n_keep * sizeof(double)) + 256);
        if (e == cudaSuccess) {
            double *w2 = reinterpret_cast<double *>(rec2 + up((size_t)n_keep * rec_bytes));
            em_flag_invert_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(flag, em->n_rows);
            em_dense_list_kernel<<<1, 1024, 0, ctx->stream>>>(flag, em->n_rows, list, d_count);
            const int g = (int)std::max<int64_t>(1, std::min<int64_t>(n_keep, (int64_t)ctx->num_sms * 8));
            em_gather_records_kernel<<<g, 256, 0, ctx->stream>>>(rec, (int64_t)rec_bytes, w_coded, list,
                                                                  n_keep, rec2, w2);
            ctx->launches += 3;
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e == cudaSuccess) {
                dev_free(ctx, rec);
                rec = rec2;
                w_coded = w2;
                n_coded = n_keep;
            } else {
                dev_free(ctx, rec2);
            }
        } else {       // no room for the compact copy: keep the records of all rows
            cudaGetLastError();
            e = cudaSuccess;
        }
    }
    if (e == cudaSuccess && worth)
        e = cudaFuncSetAttribute(two_slots ? (const void *)pick_pair_coded(em->nc)
                                           : (const void *)pick_pass_coded(em->nc, em->ld, pair_threads).fn,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)stages * rec_bytes + fixed));
    if (e == cudaSuccess && worth && two_slots && n_dense > 0)
        e = cudaFuncSetAttribute((const void *)pick_pair_accumulate(em->nc),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)em->smem_bytes);
    dev_free(ctx, tmp);
    if (e != cudaSuccess || !worth) {
        dev_free(ctx, rec);
        dev_free(ctx, dense);
        em->w_dense = nullptr;
        if (e != cudaSuccess) {
            set_error("em_pack_rows: %s", cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? MXB_ERR_NOMEM : MXB_ERR_CUDA;
        }
        return MXB_OK;
    }
    em->coded = true;
    em->pair_threads = pair_threads;
    em->rec = rec;
    em->n_coded = n_coded;
    em->rec_bytes = rec_bytes;
    em->w_coded = w_coded;
    em->dense_lin = dense;
    em->n_dense = n_dense;
    em->coded_stages = stages;
    em->coded_smem = (size_t)stages * rec_bytes + fixed;
    dev_free(ctx, em->lin);   // every pass reads the records and the dense rows from now on
    em->lin = nullptr;
    return MXB_OK;
}

static int em_create_impl(mxb_ctx *ctx, const mxb_matrix *m, const double *weights, int sharded,
                          int want_slots, mxb_em **out) {
    MXB_REQUIRE(ctx != nullptr && m != nullptr && out != nullptr, "NULL argument");
    // a rank of a row-sharded run may hold no rows at all (fewer signatures than GPUs)
    MXB_REQUIRE((m->n_rows > 0 || sharded) && m->n_cols > 0, "EM needs a non-empty matrix");
    MXB_REQUIRE(weights != nullptr || m->n_rows == 0, "weights is NULL");
    *out = nullptr;
    MXB_CUDA(cudaSetDevice(ctx->device));
    mxb_em *em = new (std::nothrow) mxb_em();
    if (!em) { set_error("out of host memory"); return MXB_ERR_NOMEM; }
    em->ctx = ctx;
    em->mat = m;
    em->n_rows = m->n_rows;
    em->n_cols = m->n_cols;
    em->ld = round_up(m->n_cols, kLdAlign);
    em->sharded = sharded != 0;

    // launch geometry
    const size_t row_bytes = (size_t)em->ld * sizeof(double);
    const size_t fixed = 2 * kPassWarps * kPassGroup * sizeof(double) + 8 * sizeof(uint64_t) + 128;
    const bool force_general = getenv("MXB_EM_FORCE_GENERAL") != nullptr;
    em->nc = (int)ceil_div(em->ld / 2, kPassThreads);
    if (!force_general && em->ld >= 1024 && em->nc <= kMaxNC && ctx->smem_optin > fixed) {
        int stages = (int)std::min<size_t>(8, (ctx->smem_optin - fixed) / row_bytes);
        if (stages >= kPassGroup + 1) {
            em->fast = true;
            em->n_stages = stages;
            em->smem_bytes = (size_t)stages * row_bytes + fixed;
            em->grid_fast = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->num_sms, em->n_rows));
        }
    }
    em->fused_tail = em->fast && em->ld <= (int64_t)kFinCtas * kFinThreads &&
                     (!(em->sharded && ctx->world > 1) || ctx->p2p_ready) &&
                     getenv("MXB_EM_SPLIT_TAIL") == nullptr;
    if (want_slots == 2 && em->fast && em->fused_tail && em->nc <= kMaxPairNC && !em->sharded)
        em->n_slots = 2;
    if (em->fast) {
        em->n_part = em->grid_fast;
        cudaError_t e = cudaFuncSetAttribute((const void *)pick_pass(em->nc),
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)em->smem_bytes);
        if (e == cudaSuccess && em->n_slots == 2)
            e = cudaFuncSetAttribute((const void *)pick_pair(em->nc),
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)em->smem_bytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%zu): %s", em->smem_bytes, cudaGetErrorString(e));
            delete em;
            return MXB_ERR_CUDA;
        }
    } else {
        em->col_blocks = (int)ceil_div(em->ld / 2, 256);
        int64_t rb = std::max<int64_t>(1, (int64_t)ctx->num_sms * 4 / em->col_blocks);
        rb = std::min<int64_t>(rb, ceil_div(em->n_rows, 16));
        em->row_blocks = (int)std::max<int64_t>(1, rb);
        em->n_part = em->row_blocks;
    }

    cudaError_t e = cudaSuccess;
#define STEP(call) do { if (e == cudaSuccess) e = (call); } while (0)
    STEP(dev_alloc(ctx, (void **)&em->lin, (size_t)em->n_rows * row_bytes));
    const size_t ns = (size_t)em->n_slots;
    {   // everything else of the session in one block: [weights][coef][lnp x2][pi x2][partials]
        // [tsum][props_in][state], each 256-byte aligned
        auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t b_rows = up((size_t)em->n_rows * sizeof(double));
        const size_t b_vec = up(ns * row_bytes);
        const size_t b_part = up(ns * (size_t)em->n_part * row_bytes);
        const size_t total = b_rows * (em->fast ? 1 : 2) + 4 * b_vec + b_part + up(row_bytes) +
                             up((size_t)em->n_cols * sizeof(double)) + up(ns * sizeof(EmState));
        STEP(dev_alloc(ctx, (void **)&em->small, total));
        if (e == cudaSuccess) {
            unsigned char *p = em->small;
            em->weights = (double *)p; p += b_rows;
            if (!em->fast) { em->coef = (double *)p; p += b_rows; }
            for (int i = 0; i < 2; ++i) { em->lnp[i] = (double *)p; p += b_vec; }
            for (int i = 0; i < 2; ++i) { em->pi[i] = (double *)p; p += b_vec; }
            em->partials = (double *)p; p += b_part;
            em->tsum = (double *)p; p += up(row_bytes);
            em->props_in = (double *)p; p += up((size_t)em->n_cols * sizeof(double));
            em->state = (EmState *)p;
        }
    }
    for (int i = 0; i < 2; ++i)
        STEP(cudaEventCreateWithFlags(&em->poll_ev[i], cudaEventDisableTiming));
    STEP(cudaMemsetAsync(em->state, 0, ns * sizeof(EmState), ctx->stream));
    if (e == cudaSuccess) {
        static_assert(2 * kMaxSlots * sizeof(EmState) <= kPinnedScratchBytes, "pinned scratch");
        em->host_state = (EmState *)pinned_scratch(ctx);
        if (!em->host_state) e = cudaErrorMemoryAllocation;
    }
    if (em->n_rows > 0)
        STEP(cudaMemcpyAsync(em->weights, weights, em->n_rows * sizeof(double),
                             cudaMemcpyHostToDevice, ctx->stream));
    if (e == cudaSuccess && em->n_rows > 0) {
        const int grid = (int)std::min<int64_t>(em->n_rows, (int64_t)ctx->num_sms * 8);
        to_linear_kernel<<<grid, 256, 0, ctx->stream>>>(m->data, em->n_rows, em->n_cols, em->ld,
                                                        em->lin);
        ctx->launches++;
        STEP(cudaGetLastError());
    }
    STEP(cudaStreamSynchronize(ctx->stream));
#undef STEP
    if (e != cudaSuccess) {
        set_error("mxb_em_create (%lld x %lld): %s", (long long)em->n_rows,
                  (long long)em->n_cols, cudaGetErrorString(e));
        mxb_em_destroy(em);
        return e == cudaErrorMemoryAllocation ? MXB_ERR_NOMEM : MXB_ERR_CUDA;
    }
    const int rc = em_pack_rows(em);
    if (rc != MXB_OK) {
        mxb_em_destroy(em);
        return rc;
    }
    *out = em;
    return MXB_OK;
}

}  // namespace mxb

extern "C" {

int mxb_em_create(mxb_ctx *ctx, const mxb_matrix *m, const double *weights, int sharded,
                  mxb_em **out) {
    return em_create_impl(ctx, m, weights, sharded, 1, out);
}

int mxb_em_pass_bytes(const mxb_em *em, int64_t *bytes_per_pass, int64_t *n_dense_rows) {
    MXB_REQUIRE(em != nullptr, "NULL argument");
    const int64_t row_bytes = em->ld * (int64_t)sizeof(double);
    if (bytes_per_pass)
        *bytes_per_pass = em->coded ? em->n_coded * (int64_t)em->rec_bytes + em->n_dense * row_bytes
                                    : em->n_rows * row_bytes;
    if (n_dense_rows) *n_dense_rows = em->coded ? em->n_dense : -1;
    return MXB_OK;
}

int mxb_em_set_lnprops(mxb_em *em, const double *lnprops) {
    MXB_REQUIRE(em != nullptr && lnprops != nullptr, "NULL argument");
    mxb_ctx *ctx = em->ctx;
    MXB_CUDA(cudaSetDevice(ctx->device));
    MXB_CUDA(cudaMemcpyAsync(em->props_in, lnprops, em->n_cols * sizeof(double),
                             cudaMemcpyHostToDevice, ctx->stream));
    em_set_props_kernel<<<(int)ceil_div(em->ld, 256), 256, 0, ctx->stream>>>(
        em->props_in, em->n_cols, em->ld, em->lnp[0], em->pi[0], em->lnp[1], em->pi[1]);
    ctx->launches++;
    MXB_CUDA(cudaGetLastError());
    MXB_CUDA(cudaStreamSynchronize(ctx->stream));  // lnprops may be a temporary
    MXB_TRY(reset_state(em, 0, 0.0));
    em->zero_iter = true;
    return MXB_OK;
}

int mxb_em_iterate(mxb_em *em, int64_t max_iter, double tol, int64_t *iters_out,
                   int32_t *converged_out) {
    MXB_REQUIRE(em != nullptr, "em is NULL");
    mxb_ctx *ctx = em->ctx;
    MXB_CUDA(cudaSetDevice(ctx->device));
    if (iters_out) *iters_out = 0;
    if (converged_out) *converged_out = 0;
    if (max_iter <= 0) {  // em.py:126 loop body never runs; :141-143 hands back the init
        em->zero_iter = true;
        return MXB_OK;
    }
    em->zero_iter = false;
    // keep `cur` as set_lnprops left it (0), restart counters
    MXB_TRY(reset_state(em, max_iter, tol));

    // Chunks of iterations are enqueued ahead of the host; a finished run turns
    // the remaining launches into early-exit no-ops.  Two chunks in flight.
    const int64_t kChunk = 16;
    int64_t enqueued = 0;
    int slot = 0;
    bool pending[2] = {false, false};
    EmState fin;
    memset(&fin, 0, sizeof(fin));
    bool finished = false;
    while (!finished) {
        if (enqueued < max_iter) {
            const int64_t n = std::min<int64_t>(kChunk, max_iter - enqueued);
            for (int64_t i = 0; i < n; ++i) MXB_TRY(enqueue_iteration(em));
            enqueued += n;
        }
        MXB_CUDA(cudaMemcpyAsync(&em->host_state[slot], em->state, sizeof(EmState),
                                 cudaMemcpyDeviceToHost, ctx->stream));
        MXB_CUDA(cudaEventRecord(em->poll_ev[slot], ctx->stream));
        pending[slot] = true;
        const int other = slot ^ 1;
        // wait for the older chunk (or this one when nothing else can be queued)
        const int wait_slot = pending[other] ? other : slot;
        const bool must_wait = pending[other] || enqueued >= max_iter;
        if (must_wait) {
            MXB_CUDA(cudaEventSynchronize(em->poll_ev[wait_slot]));
            pending[wait_slot] = false;
            if (em->host_state[wait_slot].done) {
                fin = em->host_state[wait_slot];
                finished = true;
            } else if (wait_slot == slot && enqueued >= max_iter) {
                // cannot happen: max_iter iterations always set done
                set_error("mxb_em_iterate: run did not terminate");
                return MXB_ERR_CUDA;
            }
        }
        slot = other;
    }
    MXB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (iters_out) *iters_out = fin.iters;
    if (converged_out) *converged_out = (fin.done == 1);
    if (fin.done == 3) {
        set_error("EM: a peer rank did not deliver its column sums within %llu s",
                  (unsigned long long)(kP2PTimeoutNs / 1000000000ull));
        return MXB_ERR_CUDA;
    }
    if (fin.bad) {
        set_error("EM: the mixture likelihood of %d row group(s) underflowed to 0 "
                  "(proportions left the fp64 range)", fin.bad);
        return MXB_ERR_RANGE;
    }
    return MXB_OK;
}

int mxb_em_iterate_fixed(mxb_em *em, int64_t n_iter, float *elapsed_ms, float *pass_ms) {
    MXB_REQUIRE(em != nullptr && n_iter >= 1 && n_iter <= 100000, "bad argument");
    mxb_ctx *ctx = em->ctx;
    MXB_CUDA(cudaSetDevice(ctx->device));
    em->zero_iter = false;
    MXB_TRY(reset_state(em, (long long)1 << 60, -1.0));
    std::vector<cudaEvent_t> evs;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    MXB_CUDA(cudaEventCreate(&e0));
    MXB_CUDA(cudaEventCreate(&e1));
    if (pass_ms) {
        evs.resize((size_t)n_iter * 2);
        for (auto &e : evs) MXB_CUDA(cudaEventCreate(&e));
    }
    MXB_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int64_t i = 0; i < n_iter; ++i) {
        if (pass_ms) MXB_TRY(enqueue_iteration(em, evs[2 * i], evs[2 * i + 1]));
        else MXB_TRY(enqueue_iteration(em));
    }
    MXB_CUDA(cudaEventRecord(e1, ctx->stream));
    MXB_CUDA(cudaMemcpyAsync(&em->host_state[0], em->state, sizeof(EmState),
                             cudaMemcpyDeviceToHost, ctx->stream));
    MXB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (em->host_state[0].done == 3) {
        set_error("EM: a peer rank did not deliver its column sums in time");
        return MXB_ERR_CUDA;
    }
    if (elapsed_ms) MXB_CUDA(cudaEventElapsedTime(elapsed_ms, e0, e1));
    if (pass_ms) {
        double total = 0.0;
        for (int64_t i = 0; i < n_iter; ++i) {
            float ms = 0.f;
            MXB_CUDA(cudaEventElapsedTime(&ms, evs[2 * i], evs[2 * i + 1]));
            total += ms;
        }
        *pass_ms = (float)total;
        for (auto &e : evs) cudaEventDestroy(e);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MXB_OK;
}

int mxb_em_get_lnprops(mxb_em *em, int which, double *out) {
    MXB_REQUIRE(em != nullptr && out != nullptr && (which == 0 || which == 1), "bad argument");
    mxb_ctx *ctx = em->ctx;
    MXB_CUDA(cudaSetDevice(ctx->device));
    EmState st;
    MXB_CUDA(cudaMemcpyAsync(&st, em->state, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    MXB_CUDA(cudaStreamSynchronize(ctx->stream));
    // after a finished run: lnp[cur] = previous, lnp[1-cur] = latest.
    int idx = (which == 0) ? 1 - st.cur : st.cur;
    if (st.done == 0) idx = 1 - idx;  // iterate_fixed: buffers were swapped after the last step
    if (em->zero_iter) idx = st.cur;  // no iteration ran: only the init exists
    MXB_CUDA(cudaMemcpyAsync(out, em->lnp[idx], em->n_cols * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    MXB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MXB_OK;
}

int mxb_em_read_mix(mxb_em *em, mxb_matrix *dst, int mode, double sub_log) {
    MXB_REQUIRE(em != nullptr && dst != nullptr && (mode == 0 || mode == 1), "bad argument");
    MXB_REQUIRE(dst->n_rows == em->n_rows && dst->n_cols == em->n_cols, "shape mismatch");
    mxb_ctx *ctx = em->ctx;
    MXB_CUDA(cudaSetDevice(ctx->device));
    EmState st;
    MXB_CUDA(cudaMemcpyAsync(&st, em->state, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    MXB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (em->n_rows == 0) return MXB_OK;
    const int grid = (int)std::min<int64_t>(em->n_rows, (int64_t)ctx->num_sms * 8);
    read_mix_kernel<<<grid, kMixThreads, 0, ctx->stream>>>(em->mat->data, em->n_rows, em->n_cols,
                                                           em->lnp[st.cur], dst->data, mode,
                                                           sub_log);
    ctx->launches++;
    MXB_CUDA(cudaGetLastError());
    return MXB_OK;
}

int mxb_matrix_fold_ranks(mxb_ctx *ctx, mxb_matrix *m, double sub_log) {
    MXB_REQUIRE(ctx != nullptr && m != nullptr, "NULL argument");
    MXB_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = m->n_rows * m->n_cols;
    if (n == 0) return MXB_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 16);
    double *mx = nullptr;
    MXB_CUDA(cudaMalloc(&mx, n * sizeof(double)));
    int rc = MXB_OK;
    cudaError_t e = cudaMemcpyAsync(mx, m->data, n * sizeof(double), cudaMemcpyDeviceToDevice,
                                    ctx->stream);
    if (e != cudaSuccess) rc = MXB_ERR_CUDA;
    if (rc == MXB_OK) rc = nccl_allreduce_f64(ctx, mx, n, 1);
    if (rc == MXB_OK) {
        fold_exp_kernel<<<grid, 256, 0, ctx->stream>>>(m->data, mx, n);
        ctx->launches++;
        rc = nccl_allreduce_f64(ctx, m->data, n, 0);
    }
    if (rc == MXB_OK) {
        fold_log_kernel<<<grid, 256, 0, ctx->stream>>>(m->data, mx, n, sub_log);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = MXB_ERR_CUDA;
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == MXB_OK) rc = MXB_ERR_CUDA;
    cudaFree(mx);
    if (rc == MXB_ERR_CUDA && mxb_last_error()[0] == 0) set_error("mxb_matrix_fold_ranks: CUDA failure");
    return rc;
}

}  // extern "C"

namespace mxb {

// run_em's restart loop (em.py:117-156) on a two-slot session: two restarts iterate per
// read of L; a slot whose restart has finished is handed the next one while the other
// keeps going.  The host polls the control blocks two chunks behind the device, exactly
// like mxb_em_iterate does for one restart.  Results are combined as the reference
// combines them: final log-proportions summed in restart order (em.py:145-155), read
// matrices folded with logaddexp (em.py:156; in completion order).
static int run_em_batched(mxb_em *em, const double *init_lnprops, int32_t n_multi,
                          int64_t max_iter, double tol, bool raw, mxb_matrix *mix,
                          double *props_out, int64_t *iters_out, int32_t *converged_out) {
    mxb_ctx *ctx = em->ctx;
    cudaStream_t s = ctx->stream;
    const int64_t h = em->n_cols, ld = em->ld;
    // device: [inits n_multi x h][final log-proportions n_multi x h]; the finals are read
    // back in one copy at the end (no pinned host buffer, no blocking copy mid-run)
    const size_t vec_bytes = (size_t)n_multi * h * sizeof(double);
    double *d_inits = nullptr;
    if (dev_alloc(ctx, (void **)&d_inits, 2 * vec_bytes) != cudaSuccess) {
        set_error("run_em: device allocation of %zu bytes failed", 2 * vec_bytes);
        cudaGetLastError();
        return MXB_ERR_NOMEM;
    }
    double *d_fin = d_inits + (size_t)n_multi * h;
    std::vector<double> h_fin;
    try {
        h_fin.resize((size_t)n_multi * h);
    } catch (const std::bad_alloc &) {
        dev_free(ctx, d_inits);
        set_error("run_em: out of host memory");
        return MXB_ERR_NOMEM;
    }
    int rc = copy_h2d(ctx, d_inits, init_lnprops, vec_bytes);

    int slot_restart[kMaxSlots] = {-1, -1};
    int slot_gen[kMaxSlots] = {0, 0};
    int poll_gen[2][kMaxSlots] = {{0, 0}, {0, 0}};
    int32_t next = 0, finished = 0, folded = 0;
    auto start_slot = [&](int sl) {
        const int32_t idx = next++;
        slot_restart[sl] = idx;
        slot_gen[sl]++;
        em_set_props_kernel<<<(int)ceil_div(ld, 256), 256, 0, s>>>(
            d_inits + (size_t)idx * h, h, ld, em->lnp[0] + sl * ld, em->pi[0] + sl * ld,
            em->lnp[1] + sl * ld, em->pi[1] + sl * ld);
        em_reset_state_kernel<<<1, 1, 0, s>>>(em->state + sl, (long long)max_iter, tol);
        ctx->launches += 2;
    };
    for (int sl = 0; sl < em->n_slots && rc == MXB_OK; ++sl)
        if (next < n_multi) start_slot(sl);

    const int64_t kChunk = 16;
    bool pending[2] = {false, false};
    int which = 0;
    const int grid_mix = (int)std::min<int64_t>(em->n_rows, (int64_t)ctx->num_sms * 8);
    while (rc == MXB_OK && finished < n_multi) {
        for (int64_t i = 0; i < kChunk && rc == MXB_OK; ++i) rc = enqueue_iteration(em);
        if (rc != MXB_OK) break;
        if (cudaMemcpyAsync(&em->host_state[which * kMaxSlots], em->state,
                            em->n_slots * sizeof(EmState), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaEventRecord(em->poll_ev[which], s) != cudaSuccess) {
            set_error("run_em: poll enqueue failed");
            rc = MXB_ERR_CUDA;
            break;
        }
        for (int sl = 0; sl < kMaxSlots; ++sl) poll_gen[which][sl] = slot_gen[sl];
        pending[which] = true;
        const int other = which ^ 1;
        if (pending[other]) {
            if (cudaEventSynchronize(em->poll_ev[other]) != cudaSuccess) {
                set_error("run_em: poll wait failed");
                rc = MXB_ERR_CUDA;
                break;
            }
            pending[other] = false;
            for (int sl = 0; sl < em->n_slots && rc == MXB_OK; ++sl) {
                const EmState st = em->host_state[other * kMaxSlots + sl];
                if (slot_restart[sl] < 0 || poll_gen[other][sl] != slot_gen[sl] || !st.done) continue;
                const int32_t idx = slot_restart[sl];
                if (st.bad) {
                    set_error("EM: the mixture likelihood of some rows underflowed to 0 in "
                              "restart %d (proportions left the fp64 range)", (int)idx);
                    rc = MXB_ERR_RANGE;
                    break;
                }
                if (iters_out) iters_out[idx] = st.iters;
                if (converged_out) converged_out[idx] = (st.done == 1);
                // lnp[cur] = proportions before the last step, lnp[1-cur] = after it
                if (cudaMemcpyAsync(d_fin + (size_t)idx * h, em->lnp[1 - st.cur] + sl * ld,
                                    h * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
                    set_error("run_em: result copy failed");
                    rc = MXB_ERR_CUDA;
                    break;
                }
                ++finished;
                if (mix) {
                    const bool last = finished == n_multi;
                    const double sub = (last && n_multi > 1 && !raw) ? log((double)n_multi) : 0.0;
                    read_mix_kernel<<<grid_mix, kMixThreads, 0, s>>>(
                        em->mat->data, em->n_rows, em->n_cols, em->lnp[st.cur] + sl * ld, mix->data,
                        folded == 0 ? 0 : 1, sub);
                    ctx->launches++;
                    ++folded;
                }
                slot_restart[sl] = -1;
                if (next < n_multi) start_slot(sl);
            }
        }
        which = other;
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == MXB_OK) {
        set_error("run_em: stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MXB_ERR_CUDA;
    }
    if (rc == MXB_OK) rc = copy_d2h(ctx, h_fin.data(), d_fin, vec_bytes);
    if (rc == MXB_OK) {
        for (int64_t j = 0; j < h; ++j) {
            double v = h_fin[j];
            for (int32_t i = 1; i < n_multi; ++i) v += h_fin[(size_t)i * h + j];  // em.py:155
            if (!raw) {
                if (n_multi > 1) v /= (double)n_multi;  // em.py:158-160
                v = exp(v);                             // em.py:163
            }
            props_out[j] = v;
        }
    }
    dev_free(ctx, d_inits);
    return rc;
}

}  // namespace mxb

extern "C" {

int mxb_run_em_dev(mxb_ctx *ctx, const mxb_matrix *m, const double *weights,
                   const double *init_lnprops, int32_t n_multi, int64_t max_iter, double tol,
                   int32_t flags, double *props_out, double *read_mix_out,
                   mxb_matrix **read_mix_dev, int64_t *iters_out, int32_t *converged_out) {
    MXB_REQUIRE(ctx != nullptr && m != nullptr, "NULL handle");
    MXB_REQUIRE(init_lnprops != nullptr && props_out != nullptr, "NULL buffer");
    MXB_REQUIRE(n_multi >= 1, "n_multi must be >= 1");
    if (read_mix_dev) *read_mix_dev = nullptr;
    const int64_t h = m->n_cols;
    const bool raw = (flags & MXB_EM_RAW) != 0;
    const bool want_mix = read_mix_out != nullptr || read_mix_dev != nullptr;

    mxb_em *em = nullptr;
    mxb_matrix *mix = nullptr;
    std::vector<double> acc((size_t)h, 0.0), cur((size_t)h);
    StageTimer tm(ctx->stream);
    // the caller's result buffer is usually fresh pageable memory: fault its pages in
    // from background threads while the GPU iterates
    Prefault fault_mix;
    const bool sharded = (flags & MXB_EM_SHARDED) != 0;
    const int want_slots = (n_multi >= 2 && max_iter > 0 && !sharded &&
                            getenv("MXB_EM_NO_BATCH") == nullptr) ? 2 : 1;
    int rc = em_create_impl(ctx, m, weights, sharded, want_slots, &em);
    tm.mark("em_create (alloc+to_linear)");
    if (rc == MXB_OK && want_mix) rc = mxb_matrix_alloc(ctx, m->n_rows, h, &mix);
    tm.mark("alloc read_mix");
    // (started after the device allocations: page faults and cudaMalloc contend for the
    // process's mmap lock)
    if (rc == MXB_OK && read_mix_out)
        fault_mix.start(read_mix_out, (size_t)m->n_rows * (size_t)h * sizeof(double));
    const bool batched = rc == MXB_OK && em->n_slots == 2;
    if (batched) {
        rc = run_em_batched(em, init_lnprops, n_multi, max_iter, tol, raw, mix, props_out,
                            iters_out, converged_out);
        tm.mark("iterate (two restarts per pass)");
    }
    for (int32_t i = 0; rc == MXB_OK && !batched && i < n_multi; ++i) {
        int64_t iters = 0;
        int32_t conv = 0;
        rc = mxb_em_set_lnprops(em, init_lnprops + (size_t)i * h);
        if (rc == MXB_OK) rc = mxb_em_iterate(em, max_iter, tol, &iters, &conv);
        tm.mark("iterate");
        if (iters_out) iters_out[i] = iters;
        if (converged_out) converged_out[i] = conv;
        if (rc == MXB_OK) rc = mxb_em_get_lnprops(em, 0, cur.data());
        if (rc == MXB_OK) {
            // em.py:145-155: first run stored, later runs added in place
            if (i == 0) acc = cur;
            else for (int64_t j = 0; j < h; ++j) acc[j] += cur[j];
        }
        if (rc == MXB_OK && want_mix) {
            const bool last = (i == n_multi - 1);
            const double sub = (last && n_multi > 1 && !raw) ? log((double)n_multi) : 0.0;
            rc = mxb_em_read_mix(em, mix, i == 0 ? 0 : 1, sub);
            tm.mark("read_mix kernel");
        }
    }
    if (rc == MXB_OK) {
        for (int64_t j = 0; j < h && !batched; ++j) {
            double v = acc[j];
            if (!raw) {
                if (n_multi > 1) v /= (double)n_multi;  // em.py:158-160
                v = exp(v);                             // em.py:163
            }
            props_out[j] = v;
        }
        if (read_mix_out) {
            fault_mix.join();
            rc = mxb_matrix_download(ctx, mix, read_mix_out);
            tm.mark("download read_mix");
        }
        else if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            set_error("mxb_run_em: stream sync failed");
            rc = MXB_ERR_CUDA;
        }
    }
    mxb_em_destroy(em);
    if (rc == MXB_OK && read_mix_dev) *read_mix_dev = mix;
    else mxb_matrix_destroy(mix);
    tm.mark("free");
    return rc;
}

int mxb_run_em(mxb_ctx *ctx, const double *read_hap_mat, const double *weights, int64_t n_rows,
               int64_t n_cols, const double *init_lnprops, int32_t n_multi, int64_t max_iter,
               double tol, int32_t flags, double *props_out, double *read_mix_out,
               int64_t *iters_out, int32_t *converged_out) {
    mxb_matrix *m = nullptr;
    MXB_TRY(mxb_matrix_upload(ctx, read_hap_mat, n_rows, n_cols, &m));
    int rc = mxb_run_em_dev(ctx, m, weights, init_lnprops, n_multi, max_iter, tol, flags,
                            props_out, read_mix_out, nullptr, iters_out, converged_out);
    mxb_matrix_destroy(m);
    return rc;
}

int mxb_em_step(mxb_ctx *ctx, const double *read_hap_mat, const double *weights,
                const double *ln_props, int64_t n_rows, int64_t n_cols, double *read_mix_out,
                double *new_props_out) {
    MXB_REQUIRE(read_mix_out != nullptr && new_props_out != nullptr && ln_props != nullptr,
                "NULL buffer");
    mxb_matrix *m = nullptr, *mix = nullptr;
    mxb_em *em = nullptr;
    MXB_TRY(mxb_matrix_upload(ctx, read_hap_mat, n_rows, n_cols, &m));
    int rc = mxb_em_create(ctx, m, weights, 0, &em);
    if (rc == MXB_OK) rc = mxb_matrix_alloc(ctx, n_rows, n_cols, &mix);
    if (rc == MXB_OK) rc = mxb_em_set_lnprops(em, ln_props);
    // one iteration, never "converged": afterwards previous = ln_props, latest = new
    if (rc == MXB_OK) rc = mxb_em_iterate(em, 1, -1.0, nullptr, nullptr);
    if (rc == MXB_OK) rc = mxb_em_get_lnprops(em, 0, new_props_out);
    if (rc == MXB_OK) rc = mxb_em_read_mix(em, mix, 0, 0.0);
    if (rc == MXB_OK) rc = mxb_matrix_download(ctx, mix, read_mix_out);
    mxb_em_destroy(em);
    mxb_matrix_destroy(mix);
    mxb_matrix_destroy(m);
    return rc;
}

}  // extern "C"
