// Kernel 1: read-signature x haplotype log-likelihood matrix.
//
// Replaces the N x H x K Python loop of build_em_matrix
// (reference mixemt/preprocess.py:177-198) and the per-cell dictionary walk of
// HapVarBaseMatrix._prob / prob_for_vars (preprocess.py:69-96).
//
// Haplotype tables.  For every variant position p and symbol code a the dense
// table bits[p][a][w] holds one bit per haplotype column: bit (j & 31) of word
// (j >> 5) is set iff haplotype j *expects* symbol a at p (its marker there,
// or the reference base when it carries none: preprocess.py:75-83).  A 32-bit
// word is a ready-made ballot over 32 haplotypes for one observed
// (position, base); plane n_sym is all zero and serves bases that can never
// match.  99 % of the expected bases are the reference base, so the table is
// also kept in *sparse deviation* form: for plane (p, a) the list of
// (word index g, D) with D = bits[p][a][g] XOR (a == ref[p] ? all ones : 0)
// != 0, i.e. the haplotypes whose match/mismatch outcome differs from that of
// a marker-free haplotype.  Build 17: 4070 x 5 x 172 dense words (14 MB) against
// ~0.3 M sparse entries.
//
// Bit-exactness.  A cell is the fp64 sum, in signature order, of hit[p] on a
// match and miss[p] on a mismatch -- the same values in the same order as the
// reference's `total += math.log(...)` (preprocess.py:92-95).  fp64 addition is
// not associative, so the K terms of a cell cannot be re-grouped; what can be
// shared is whole chains: two haplotypes with the same match pattern over the
// row's K positions get the same sum, and a chain whose first deviation from
// the marker-free pattern is at k0 may start from the marker-free prefix sum
// P[k0] (the same additions in the same order).  A 300 bp fragment separates
// the 5408 Build-17 haplotypes into only ~450 (32-column group, pattern)
// classes, so the class kernel evaluates ~12x fewer dependent-add chains than
// cells, each about half as long.
//
// build_matrix_kernel (class kernel), one CTA (512 threads) per row at a time:
//   0. stage the row: per observation (base term, deviating term), the range
//      of its sparse deviation list, whether the marker-free outcome is a match;
//   1. one thread forms the marker-free prefix sums P[0..K] (and match counts)
//      while the other warps scatter the row's deviation entries into
//      per-group lists in shared memory (shared-memory atomics);
//   2. one warp per 32-column group: sort the group's (k, D) entries by k, give
//      every lane its pattern over them, find the distinct patterns with
//      match.any and allocate one chain item per distinct non-empty pattern;
//   3. one thread per chain item: start at P[k0], add the remaining terms in
//      order, taking the deviating term where the pattern says so;
//   4. every cell looks up the value of its class and the row is written with
//      coalesced 16-byte stores.
// Rows or groups that exceed the shared-memory budgets (K > 512 observations,
// more than 16 deviating positions in a group, more than 2048 chains in a row)
// take the dense path: build_dense_row / the per-group dense loop walk the
// dense bitset table cell by cell, exactly like build_matrix_dense_kernel, the
// plain kernel that is also used when H is too large for the class kernel's
// shared-memory layout (or MXB_BUILD_DENSE=1 is set).
#include <new>
#include <vector>

#include "common.cuh"

namespace mxb {

constexpr int kBuildThreads = 512;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kBuildCols = 4;        // dense path: haplotype columns per thread and group
constexpr int kBuildObsChunk = 512;  // observations staged per pass (dense path), KMAX (class path)
constexpr int kGroupCap = 16;        // deviating observations kept per 32-column group
constexpr int kMaxItems = 2048;      // chain items per row
constexpr unsigned kNoItem = 0xFFFFu;

// ---- dense path ---------------------------------------------------------------
// One CTA walks one row; a thread owns 4 haplotype columns per 2048-column
// group (same bit lane, 4 different words), so every bitset load is
// warp-uniform and the four accumulation chains are independent.
__device__ __forceinline__ void build_dense_row(
        const uint32_t *__restrict__ bits, const double2 *__restrict__ hitmiss, int n_planes,
        int n_words, int n_hap, int64_t k0, int64_t k1, const int32_t *__restrict__ pos_idx,
        const uint8_t *__restrict__ base_code, double *__restrict__ out_row,
        int32_t *__restrict__ match_row, double2 *s_hm, uint32_t *s_off) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int kGroup = kBuildThreads * kBuildCols;
    constexpr int kWordsPerSlot = kBuildThreads / 32;

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
    __syncthreads();
}

__global__ void __launch_bounds__(kBuildThreads)
build_matrix_dense_kernel(const uint32_t *__restrict__ bits, const double2 *__restrict__ hitmiss,
                          int n_planes, int n_words, int n_hap, int64_t n_rows,
                          const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ pos_idx,
                          const uint8_t *__restrict__ base_code, double *__restrict__ out,
                          int32_t *__restrict__ match_out) {
    __shared__ double2 s_hm[kBuildObsChunk];
    __shared__ uint32_t s_off[kBuildObsChunk];
    for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        build_dense_row(bits, hitmiss, n_planes, n_words, n_hap, row_ptr[row], row_ptr[row + 1],
                        pos_idx, base_code, out + row * (int64_t)n_hap,
                        match_out ? match_out + row * (int64_t)n_hap : nullptr, s_hm, s_off);
    }
}

// ---- class kernel ---------------------------------------------------------------
struct BuildTables {
    const uint32_t *bits;      // dense [n_pos][n_planes][n_words]
    const double2 *hitmiss;    // [n_pos]
    const uint8_t *ref_code;   // [n_pos]
    const int32_t *dev_ptr;    // [n_pos * n_planes + 1]
    const uint2 *dev_ent;      // {word index, D}
    int n_planes, n_words, n_hap, n_groups;
};

// Shared-memory layout of the class kernel (bytes), shared by host and device.
struct BuildSmem {
    size_t term, prefix, range, off, pcnt, gcnt, glist, cell, item, icnt, total;
    __host__ __device__ BuildSmem(int n_groups, bool counts) {
        size_t o = 0;
        term = o;   o += sizeof(double2) * kBuildObsChunk;          // (base, deviating) terms
        prefix = o; o += sizeof(double) * (kBuildObsChunk + 1);     // marker-free prefix sums
        glist = o;  o += sizeof(uint2) * (size_t)kGroupCap * n_groups;  // per-group (k, D)
        item = o;   o += sizeof(uint2) * kMaxItems;                 // (pattern, group) -> value
        range = o;  o += sizeof(int2) * kBuildObsChunk;             // sparse list range per k
        off = o;    o += sizeof(uint32_t) * kBuildObsChunk;         // dense plane offset per k
        gcnt = o;   o += sizeof(int) * (size_t)n_groups;            // entries per group
        icnt = o;   o += counts ? sizeof(int) * kMaxItems : 0;      // match count per item
        pcnt = o;   o += sizeof(uint16_t) * (kBuildObsChunk + 2);   // marker-free match prefix
        cell = o;   o += sizeof(uint16_t) * (size_t)n_groups * 32;  // item of every cell
        total = (o + 15) & ~(size_t)15;
    }
};

template <bool kCounts>
__global__ void __launch_bounds__(kBuildThreads)
build_matrix_kernel(BuildTables tb, int64_t n_rows, const int64_t *__restrict__ row_ptr,
                    const int32_t *__restrict__ pos_idx, const uint8_t *__restrict__ base_code,
                    double *__restrict__ out, int32_t *__restrict__ match_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_nitems;
    __shared__ int s_overflow;
    const BuildSmem L(tb.n_groups, kCounts);
    double2 *s_term = reinterpret_cast<double2 *>(smem + L.term);
    double *s_prefix = reinterpret_cast<double *>(smem + L.prefix);
    uint2 *s_glist = reinterpret_cast<uint2 *>(smem + L.glist);
    uint2 *s_item = reinterpret_cast<uint2 *>(smem + L.item);
    double *s_val = reinterpret_cast<double *>(smem + L.item);  // overwrites the item
    int2 *s_range = reinterpret_cast<int2 *>(smem + L.range);
    uint32_t *s_off = reinterpret_cast<uint32_t *>(smem + L.off);
    int *s_gcnt = reinterpret_cast<int *>(smem + L.gcnt);
    int *s_icnt = reinterpret_cast<int *>(smem + L.icnt);
    uint16_t *s_pcnt = reinterpret_cast<uint16_t *>(smem + L.pcnt);
    uint16_t *s_cell = reinterpret_cast<uint16_t *>(smem + L.cell);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int n_hap = tb.n_hap;
    const int n_groups = tb.n_groups;

    for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const int64_t k0 = row_ptr[row];
        const int64_t k1 = row_ptr[row + 1];
        double *out_row = out + row * (int64_t)n_hap;
        int32_t *match_row = kCounts ? match_out + row * (int64_t)n_hap : nullptr;
        if (k1 - k0 > kBuildObsChunk) {  // block-uniform: long rows go dense
            build_dense_row(tb.bits, tb.hitmiss, tb.n_planes, tb.n_words, n_hap, k0, k1, pos_idx,
                            base_code, out_row, match_row, s_term, s_off);
            continue;
        }
        const int n_obs = (int)(k1 - k0);

        // ---- 0. stage the row -------------------------------------------------------
        for (int k = tid; k < n_obs; k += kBuildThreads) {
            const int p = pos_idx[k0 + k];
            int c = base_code[k0 + k];
            if (c >= tb.n_planes - 1) c = tb.n_planes - 1;
            const int plane = p * tb.n_planes + c;
            const double2 hm = tb.hitmiss[p];
            const bool base_match = (c == (int)tb.ref_code[p]);
            s_term[k] = base_match ? hm : make_double2(hm.y, hm.x);
            s_range[k] = make_int2(tb.dev_ptr[plane], tb.dev_ptr[plane + 1]);
            s_off[k] = (uint32_t)(plane * tb.n_words);
            // the match flag rides in s_pcnt[k + 1] until the prefix pass
            s_pcnt[k + 1] = base_match ? 1 : 0;
        }
        for (int g = tid; g < n_groups; g += kBuildThreads) s_gcnt[g] = 0;
        if (tid == 0) { s_nitems = 1; s_overflow = 0; }
        __syncthreads();

        // ---- 1. marker-free prefix sums | scatter deviations into group lists ----------
        if (warp == kBuildWarps - 1) {
            if (lane == 0) {
                double acc = 0.0;
                int cnt = 0;
                s_prefix[0] = acc;
                s_pcnt[0] = 0;
                for (int k = 0; k < n_obs; ++k) {
                    acc += s_term[k].x;
                    cnt += s_pcnt[k + 1];
                    s_prefix[k + 1] = acc;
                    s_pcnt[k + 1] = (uint16_t)cnt;
                }
            }
        } else {
            for (int k = warp; k < n_obs; k += kBuildWarps - 1) {
                const int2 r = s_range[k];
                for (int e = r.x + lane; e < r.y; e += 32) {
                    const uint2 ent = tb.dev_ent[e];
                    const int slot = atomicAdd(&s_gcnt[ent.x], 1);
                    if (slot < kGroupCap) s_glist[ent.x * kGroupCap + slot] = make_uint2(k, ent.y);
                }
            }
        }
        __syncthreads();

        // ---- 2. classes of every 32-column group ------------------------------------------
        for (int g = warp; g < n_groups; g += kBuildWarps) {
            const int a = s_gcnt[g];
            if (a == 0) {
                s_cell[g * 32 + lane] = 0;
                continue;
            }
            if (a > kGroupCap) {
                // too many deviating positions: this warp walks the dense table
                const int j = g * 32 + lane;
                double acc = 0.0;
                int cnt = 0;
                for (int k = 0; k < n_obs; ++k) {
                    const uint32_t w = __ldg(tb.bits + s_off[k] + g);
                    const uint32_t d = ((w >> lane) & 1u);   // 1 = match
                    const double2 t = s_term[k];
                    const bool bm = s_pcnt[k + 1] != s_pcnt[k];
                    acc += (d != 0) == bm ? t.x : t.y;
                    cnt += d;
                }
                if (j < n_hap) {
                    out_row[j] = acc;
                    if (kCounts) match_row[j] = cnt;
                }
                s_cell[g * 32 + lane] = (uint16_t)kNoItem;
                continue;
            }
            // lane e < a holds entry e; rank it by k (k is unique within a group)
            uint2 ent = make_uint2(0xFFFFFFFFu, 0u);
            if (lane < a) ent = s_glist[g * kGroupCap + lane];
            int rank = 0;
            uint32_t pattern = 0;
            for (int e = 0; e < a; ++e) {
                const uint32_t ke = __shfl_sync(0xffffffffu, ent.x, e);
                rank += (ke < ent.x) ? 1 : 0;
            }
            for (int e = 0; e < a; ++e) {
                const uint32_t we = __shfl_sync(0xffffffffu, ent.y, e);
                const int re = __shfl_sync(0xffffffffu, rank, e);
                pattern |= ((we >> lane) & 1u) << re;
            }
            __syncwarp();
            if (lane < a) s_glist[g * kGroupCap + rank] = ent;  // sorted by k
            const uint32_t peers = __match_any_sync(0xffffffffu, pattern);
            const int leader_lane = __ffs(peers) - 1;
            const bool leader = (lane == leader_lane) && pattern != 0u;
            const uint32_t lead_mask = __ballot_sync(0xffffffffu, leader);
            int base_idx = 0;
            if (lane == 0 && lead_mask) base_idx = atomicAdd(&s_nitems, __popc(lead_mask));
            base_idx = __shfl_sync(0xffffffffu, base_idx, 0);
            int idx = base_idx + __popc(lead_mask & ((1u << lane) - 1u));
            if (leader && idx < kMaxItems) s_item[idx] = make_uint2(pattern, (uint32_t)g);
            idx = __shfl_sync(0xffffffffu, idx, leader_lane);
            if (pattern == 0u) idx = 0;
            if (base_idx + __popc(lead_mask) > kMaxItems) {
                if (lane == 0) s_overflow = 1;
                idx = 0;
            }
            s_cell[g * 32 + lane] = (uint16_t)idx;
        }
        __syncthreads();
        if (s_overflow) {  // block-uniform: more classes than chain items
            build_dense_row(tb.bits, tb.hitmiss, tb.n_planes, tb.n_words, n_hap, k0, k1, pos_idx,
                            base_code, out_row, match_row, s_term, s_off);
            continue;
        }

        // ---- 3. one dependent-add chain per class ----------------------------------------------
        const int n_items = s_nitems;
        if (tid == 0) {
            s_val[0] = s_prefix[n_obs];
            if (kCounts) s_icnt[0] = s_pcnt[n_obs];
        }
        for (int it = 1 + tid; it < n_items; it += kBuildThreads) {
            const uint2 item = s_item[it];
            const uint2 *list = s_glist + item.y * kGroupCap;
            uint32_t rem = item.x;
            int e = __ffs(rem) - 1;
            rem >>= e;
            int nk = (int)list[e].x;
            double acc = s_prefix[nk];
            int cnt = kCounts ? (int)s_pcnt[nk] : 0;
            for (int k = nk; k < n_obs; ++k) {
                const double2 t = s_term[k];
                const bool dev = (k == nk);
                acc += dev ? t.y : t.x;
                if (kCounts) {
                    const int bm = (int)s_pcnt[k + 1] - (int)s_pcnt[k];
                    cnt += dev ? 1 - bm : bm;
                }
                if (dev) {
                    rem >>= 1;
                    if (rem) {
                        const int s = __ffs(rem) - 1;
                        rem >>= s;
                        e += s + 1;
                        nk = (int)list[e].x;
                    } else {
                        nk = -1;
                    }
                }
            }
            s_val[it] = acc;
            if (kCounts) s_icnt[it] = cnt;
        }
        __syncthreads();

        // ---- 4. write the row -----------------------------------------------------------------
        if (((n_hap & 1) == 0)) {
            for (int j = tid * 2; j < n_hap; j += kBuildThreads * 2) {
                const unsigned i0 = s_cell[j], i1 = s_cell[j + 1];
                if (i0 == kNoItem) continue;  // dense group (both cells are in it)
                *reinterpret_cast<double2 *>(out_row + j) = make_double2(s_val[i0], s_val[i1]);
                if (kCounts) *reinterpret_cast<int2 *>(match_row + j) = make_int2(s_icnt[i0], s_icnt[i1]);
            }
        } else {
            for (int j = tid; j < n_hap; j += kBuildThreads) {
                const unsigned i0 = s_cell[j];
                if (i0 == kNoItem) continue;
                out_row[j] = s_val[i0];
                if (kCounts) match_row[j] = s_icnt[i0];
            }
        }
        __syncthreads();  // the row's shared state is reused by the next row
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
    // sparse deviation lists: D = bits XOR (marker-free outcome), non-zero words only
    std::vector<int32_t> dev_ptr;
    std::vector<uint2> dev_ent;
    try {
        dev_ptr.assign((size_t)n_pos * n_planes + 1, 0);
        const int n_groups = (int)ceil_div(std::max(n_hap, 1), 32);
        for (int p = 0; p < n_pos; ++p) {
            for (int a = 0; a < n_planes; ++a) {
                const uint32_t *plane = &bits[((size_t)p * n_planes + a) * n_words];
                const bool base_match = (a == (int)ref_code[p]);
                for (int g = 0; g < n_groups; ++g) {
                    const int valid = std::min(32, n_hap - g * 32);
                    const uint32_t vmask = valid >= 32 ? 0xFFFFFFFFu
                                           : (valid <= 0 ? 0u : ((1u << valid) - 1u));
                    const uint32_t d = (plane[g] ^ (base_match ? 0xFFFFFFFFu : 0u)) & vmask;
                    if (d) dev_ent.push_back(make_uint2((uint32_t)g, d));
                }
                dev_ptr[(size_t)p * n_planes + a + 1] = (int32_t)dev_ent.size();
            }
        }
    } catch (const std::bad_alloc &) {
        set_error("out of host memory");
        return MXB_ERR_NOMEM;
    }
    MXB_CUDA(cudaSetDevice(ctx->device));
    mxb_phylo *ph = new (std::nothrow) mxb_phylo();
    if (!ph) { set_error("out of host memory"); return MXB_ERR_NOMEM; }
    ph->ctx = ctx;
    ph->n_pos = n_pos;
    ph->n_hap = n_hap;
    ph->n_sym = n_sym;
    ph->n_words = n_words;
    ph->n_dev = (int64_t)dev_ent.size();
    cudaError_t e = cudaSuccess;
    if (!bits.empty()) {
        const size_t ent_bytes = std::max<size_t>(1, dev_ent.size()) * sizeof(uint2);
        e = cudaMalloc(&ph->bits, bits.size() * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&ph->hitmiss, hm.size() * sizeof(double2));
        if (e == cudaSuccess) e = cudaMalloc(&ph->ref_code, (size_t)n_pos);
        if (e == cudaSuccess) e = cudaMalloc(&ph->dev_ptr, dev_ptr.size() * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc(&ph->dev_ent, ent_bytes);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ph->bits, bits.data(), bits.size() * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ph->hitmiss, hm.data(), hm.size() * sizeof(double2),
                                cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ph->ref_code, ref_code, (size_t)n_pos, cudaMemcpyHostToDevice,
                                ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ph->dev_ptr, dev_ptr.data(), dev_ptr.size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && !dev_ent.empty())
            e = cudaMemcpyAsync(ph->dev_ent, dev_ent.data(), dev_ent.size() * sizeof(uint2),
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
    if (ph->ref_code) cudaFree(ph->ref_code);
    if (ph->dev_ptr) cudaFree(ph->dev_ptr);
    if (ph->dev_ent) cudaFree(ph->dev_ent);
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
            // class kernel when its shared-memory layout fits, else the dense kernel
            const bool counts = d_match != nullptr;
            const int n_groups = (int)ceil_div(ph->n_hap, 32);
            const BuildSmem lay(n_groups, counts);
            const void *fn = counts ? (const void *)build_matrix_kernel<true>
                                    : (const void *)build_matrix_kernel<false>;
            bool use_class = getenv("MXB_BUILD_DENSE") == nullptr && n_groups < (int)kNoItem / 32 &&
                             lay.total + 1024 <= ctx->smem_optin;
            int per_sm = 0;
            if (use_class) {
                STEP(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)lay.total));
                STEP(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBuildThreads,
                                                                   lay.total));
                if (e == cudaSuccess && per_sm < 1) use_class = false;
            }
            STEP(cudaEventRecord(ev0, ctx->stream));
            if (e == cudaSuccess && use_class) {
                BuildTables tb;
                tb.bits = ph->bits;
                tb.hitmiss = ph->hitmiss;
                tb.ref_code = ph->ref_code;
                tb.dev_ptr = ph->dev_ptr;
                tb.dev_ent = ph->dev_ent;
                tb.n_planes = ph->n_sym + 1;
                tb.n_words = ph->n_words;
                tb.n_hap = ph->n_hap;
                tb.n_groups = n_groups;
                const int grid = (int)std::min<int64_t>(n_rows, (int64_t)ctx->num_sms * per_sm);
                if (counts)
                    build_matrix_kernel<true><<<grid, kBuildThreads, lay.total, ctx->stream>>>(
                        tb, n_rows, d_row_ptr, d_pos, d_code, m->data, d_match);
                else
                    build_matrix_kernel<false><<<grid, kBuildThreads, lay.total, ctx->stream>>>(
                        tb, n_rows, d_row_ptr, d_pos, d_code, m->data, d_match);
            } else if (e == cudaSuccess) {
                const int grid = (int)std::min<int64_t>(n_rows, (int64_t)ctx->num_sms * 4);
                build_matrix_dense_kernel<<<grid, kBuildThreads, 0, ctx->stream>>>(
                    ph->bits, ph->hitmiss, ph->n_sym + 1, ph->n_words, ph->n_hap, n_rows,
                    d_row_ptr, d_pos, d_code, m->data, d_match);
            }
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
