// Kernel 1: read-signature x haplotype log-likelihood matrix.
//
// Replaces the N x H x K Python loop of build_em_matrix
// (reference mixemt/preprocess.py:177-198) and the per-cell dictionary walk of
// HapVarBaseMatrix._prob / prob_for_vars (preprocess.py:69-96).
//
// Data layout.  For every variant position p and symbol code a the table
// bits[p][a][w] holds one bit per haplotype column: bit (j & 31) of word
// (j >> 5) is set iff haplotype j *expects* symbol a at p (its marker there,
// or the reference base when it carries none: preprocess.py:75-83).  A 32-bit
// word is therefore a ready-made ballot over 32 haplotypes for one observed
// (position, base); plane n_sym is all zero and serves bases that can never
// match (anything outside the alphabet).  Table size at Build 17:
// 4070 x 5 x 172 words = 14 MB, L2-resident.
//
// Work mapping.  One CTA walks one signature row at a time; a thread owns
// 4 haplotype columns per 1024-column group (same bit lane, 4 different
// words), so every bitset load is warp-uniform (one broadcast transaction per
// warp) and the four accumulation chains are independent.  The row's observed
// positions are staged once in shared memory as (plane offset, hit, miss).
// Each cell is the fp64 sum, in signature order, of hit[p] on a match and
// miss[p] on a mismatch -- the same values in the same order as the
// reference's `total += math.log(...)` (preprocess.py:92-95), so the result is
// bit-identical, and the int32 match count is the popcount of the cell's match
// bits.
#include <new>
#include <vector>

#include "common.cuh"

namespace mxb {

constexpr int kBuildThreads = 256;
constexpr int kBuildCols = 4;     // haplotype columns per thread and group
constexpr int kBuildObsChunk = 512;

struct ObsEntry {
    double hit;
    double miss;
};

__global__ void __launch_bounds__(kBuildThreads)
build_matrix_kernel(const uint32_t *__restrict__ bits, const double2 *__restrict__ hitmiss,
                    int n_planes, int n_words, int n_hap, int64_t n_rows,
                    const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ pos_idx,
                    const uint8_t *__restrict__ base_code, double *__restrict__ out,
                    int32_t *__restrict__ match_out) {
    __shared__ double2 s_hm[kBuildObsChunk];
    __shared__ uint32_t s_off[kBuildObsChunk];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int kGroup = kBuildThreads * kBuildCols;  // 1024 columns
    constexpr int kWordsPerSlot = kBuildThreads / 32;   // 8 words per u-slot

    for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const int64_t k0 = row_ptr[row];
        const int64_t k1 = row_ptr[row + 1];
        double *out_row = out + row * (int64_t)n_hap;
        int32_t *match_row = match_out ? match_out + row * (int64_t)n_hap : nullptr;

        for (int64_t kc = k0; kc < k1 || kc == k0; kc += kBuildObsChunk) {
            const int n_obs = (int)min((int64_t)kBuildObsChunk, k1 - kc);
            __syncthreads();  // previous users of s_hm / s_off are done
            for (int k = tid; k < n_obs; k += kBuildThreads) {
                const int p = pos_idx[kc + k];
                int c = base_code[kc + k];
                if (c >= n_planes - 1) c = n_planes - 1;  // "other": all-zero plane
                s_off[k] = (uint32_t)((p * n_planes + c) * n_words);
                s_hm[k] = hitmiss[p];
            }
            __syncthreads();
            const bool first = (kc == k0);

            for (int g0 = 0; g0 < n_hap; g0 += kGroup) {
                double acc[kBuildCols];
                int cnt[kBuildCols];
                const int word0 = (g0 >> 5) + warp;
#pragma unroll
                for (int u = 0; u < kBuildCols; ++u) {
                    const int j = g0 + u * kBuildThreads + tid;
                    acc[u] = 0.0;
                    cnt[u] = 0;
                    if (!first && j < n_hap) {
                        acc[u] = out_row[j];
                        if (match_row) cnt[u] = match_row[j];
                    }
                }
                for (int k = 0; k < n_obs; ++k) {
                    const uint32_t *plane = bits + s_off[k] + word0;
                    const double2 hm = s_hm[k];
                    uint32_t w[kBuildCols];
#pragma unroll
                    for (int u = 0; u < kBuildCols; ++u) {
                        // words past the table end belong to columns >= n_hap
                        w[u] = (word0 + u * kWordsPerSlot < n_words)
                                   ? __ldg(plane + u * kWordsPerSlot) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < kBuildCols; ++u) {
                        const bool hit = (w[u] >> lane) & 1u;
                        acc[u] += hit ? hm.x : hm.y;
                        cnt[u] += hit ? 1 : 0;
                    }
                }
#pragma unroll
                for (int u = 0; u < kBuildCols; ++u) {
                    const int j = g0 + u * kBuildThreads + tid;
                    if (j < n_hap) {
                        out_row[j] = acc[u];
                        if (match_row) match_row[j] = cnt[u];
                    }
                }
            }
            if (k1 == k0) break;  // empty row: cells are 0.0 (never produced by the reference)
        }
    }
}

}  // namespace mxb

using namespace mxb;

extern "C" {

int mxb_phylo_pack(mxb_ctx *ctx, int32_t n_pos, int32_t n_hap, int32_t n_sym,
                   const double *hit, const double *miss, const uint8_t *ref_code,
                   const int64_t *marker_ptr, const int32_t *marker_pos_idx,
                   const uint8_t *marker_code, mxb_phylo **out) {
    MXB_REQUIRE(ctx != nullptr && out != nullptr, "NULL argument");
    MXB_REQUIRE(n_pos >= 0 && n_hap >= 0 && n_sym >= 1 && n_sym <= 254, "bad sizes");
    MXB_REQUIRE(n_pos == 0 || (hit && miss && ref_code), "NULL table");
    MXB_REQUIRE(n_hap == 0 || marker_ptr, "NULL marker_ptr");
    *out = nullptr;
    const int n_planes = n_sym + 1;
    const int n_words = (int)round_up(ceil_div(std::max(n_hap, 1), 32), 4);
    if ((int64_t)n_pos * n_planes * n_words >= ((int64_t)1 << 31)) {
        set_error("mxb_phylo_pack: bitset table too large (%d x %d x %d words)",
                  n_pos, n_planes, n_words);
        return MXB_ERR_ARG;
    }
    std::vector<uint32_t> bits;
    std::vector<double2> hm;
    try {
        bits.assign((size_t)n_pos * n_planes * n_words, 0u);
        hm.resize((size_t)n_pos);
    } catch (const std::bad_alloc &) {
        set_error("out of host memory");
        return MXB_ERR_NOMEM;
    }
    for (int p = 0; p < n_pos; ++p) {
        if (ref_code[p] >= n_sym) {
            set_error("mxb_phylo_pack: ref_code[%d]=%d outside the alphabet", p, ref_code[p]);
            return MXB_ERR_ARG;
        }
        uint32_t *plane = &bits[((size_t)p * n_planes + ref_code[p]) * n_words];
        for (int j = 0; j < n_hap; ++j) plane[j >> 5] |= 1u << (j & 31);
        hm[p] = make_double2(hit[p], miss[p]);
    }
    for (int j = 0; j < n_hap; ++j) {
        for (int64_t m = marker_ptr[j]; m < marker_ptr[j + 1]; ++m) {
            const int p = marker_pos_idx[m];
            const int c = marker_code[m];
            if (p < 0 || p >= n_pos || c >= n_sym) {
                set_error("mxb_phylo_pack: marker %lld of haplotype %d out of range",
                          (long long)m, j);
                return MXB_ERR_ARG;
            }
            for (int a = 0; a < n_sym; ++a)
                bits[((size_t)p * n_planes + a) * n_words + (j >> 5)] &= ~(1u << (j & 31));
            bits[((size_t)p * n_planes + c) * n_words + (j >> 5)] |= 1u << (j & 31);
        }
    }
    MXB_CUDA(cudaSetDevice(ctx->device));
    mxb_phylo *ph = new (std::nothrow) mxb_phylo();
    if (!ph) { set_error("out of host memory"); return MXB_ERR_NOMEM; }
    ph->ctx = ctx;
    ph->n_pos = n_pos;
    ph->n_hap = n_hap;
    ph->n_sym = n_sym;
    ph->n_words = n_words;
    cudaError_t e = cudaSuccess;
    if (!bits.empty()) {
        e = cudaMalloc(&ph->bits, bits.size() * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&ph->hitmiss, hm.size() * sizeof(double2));
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ph->bits, bits.data(), bits.size() * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ph->hitmiss, hm.data(), hm.size() * sizeof(double2),
                                cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) {
        set_error("mxb_phylo_pack: %s", cudaGetErrorString(e));
        mxb_phylo_destroy(ph);
        return MXB_ERR_CUDA;
    }
    *out = ph;
    return MXB_OK;
}

int mxb_phylo_destroy(mxb_phylo *ph) {
    if (!ph) return MXB_OK;
    cudaSetDevice(ph->ctx->device);
    if (ph->bits) cudaFree(ph->bits);
    if (ph->hitmiss) cudaFree(ph->hitmiss);
    delete ph;
    return MXB_OK;
}

int mxb_build_matrix(mxb_ctx *ctx, const mxb_phylo *ph, int64_t n_rows,
                     const int64_t *row_ptr, const int32_t *pos_idx,
                     const uint8_t *base_code, double *out_host, int32_t *match_host,
                     mxb_matrix **out_dev, float *elapsed_ms) {
    MXB_REQUIRE(ctx != nullptr && ph != nullptr, "NULL handle");
    MXB_REQUIRE(n_rows >= 0, "negative n_rows");
    MXB_REQUIRE(n_rows == 0 || row_ptr != nullptr, "row_ptr is NULL");
    if (out_dev) *out_dev = nullptr;
    if (elapsed_ms) *elapsed_ms = 0.f;
    MXB_CUDA(cudaSetDevice(ctx->device));
    const int64_t n_obs = n_rows ? row_ptr[n_rows] : 0;
    MXB_REQUIRE(n_obs == 0 || (pos_idx && base_code), "NULL observation arrays");
    MXB_REQUIRE(n_obs == 0 || ph->n_pos > 0, "observations but no positions");

    mxb_matrix *m = nullptr;
    MXB_TRY(mxb_matrix_alloc(ctx, n_rows, ph->n_hap, &m));
    const size_t cells = (size_t)n_rows * (size_t)ph->n_hap;
    int64_t *d_row_ptr = nullptr;
    int32_t *d_pos = nullptr, *d_match = nullptr;
    uint8_t *d_code = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaError_t e = cudaSuccess;
    int rc = MXB_OK;
#define STEP(call) do { if (e == cudaSuccess) e = (call); } while (0)
    if (cells) {
        STEP(cudaMalloc(&d_row_ptr, (n_rows + 1) * sizeof(int64_t)));
        if (n_obs) {
            STEP(cudaMalloc(&d_pos, n_obs * sizeof(int32_t)));
            STEP(cudaMalloc(&d_code, n_obs * sizeof(uint8_t)));
        }
        if (match_host) STEP(cudaMalloc(&d_match, cells * sizeof(int32_t)));
        STEP(cudaMemcpyAsync(d_row_ptr, row_ptr, (n_rows + 1) * sizeof(int64_t),
                             cudaMemcpyHostToDevice, ctx->stream));
        if (n_obs) {
            STEP(cudaMemcpyAsync(d_pos, pos_idx, n_obs * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, ctx->stream));
            STEP(cudaMemcpyAsync(d_code, base_code, n_obs * sizeof(uint8_t),
                                 cudaMemcpyHostToDevice, ctx->stream));
        }
        STEP(cudaEventCreate(&ev0));
        STEP(cudaEventCreate(&ev1));
        if (e == cudaSuccess) {
            const int grid = (int)std::min<int64_t>(n_rows, (int64_t)ctx->num_sms * 8);
            STEP(cudaEventRecord(ev0, ctx->stream));
            build_matrix_kernel<<<grid, kBuildThreads, 0, ctx->stream>>>(
                ph->bits, ph->hitmiss, ph->n_sym + 1, ph->n_words, ph->n_hap, n_rows,
                d_row_ptr, d_pos, d_code, m->data, d_match);
            ctx->launches++;
            STEP(cudaGetLastError());
            STEP(cudaEventRecord(ev1, ctx->stream));
        }
        if (out_host)
            STEP(cudaMemcpyAsync(out_host, m->data, cells * sizeof(double),
                                 cudaMemcpyDeviceToHost, ctx->stream));
        if (match_host)
            STEP(cudaMemcpyAsync(match_host, d_match, cells * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, ctx->stream));
        STEP(cudaStreamSynchronize(ctx->stream));
        if (e == cudaSuccess && elapsed_ms) STEP(cudaEventElapsedTime(elapsed_ms, ev0, ev1));
    }
#undef STEP
    if (e != cudaSuccess) {
        set_error("mxb_build_matrix: %s", cudaGetErrorString(e));
        rc = (e == cudaErrorMemoryAllocation) ? MXB_ERR_NOMEM : MXB_ERR_CUDA;
    }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    cudaFree(d_row_ptr);
    cudaFree(d_pos);
    cudaFree(d_code);
    cudaFree(d_match);
    if (rc != MXB_OK || !out_dev) mxb_matrix_destroy(m);
    else *out_dev = m;
    return rc;
}

}  // extern "C"
