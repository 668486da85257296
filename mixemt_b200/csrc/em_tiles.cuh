// Class tiles: the EM pass over column classes instead of columns (csrc/em.cu: em_pack_tiles).
//
// A fragment covers ~78 of the 4070 variant positions, so over a batch of neighbouring
// signature rows (the rows arrive string-sorted, preprocess.py:219: neighbours start at the
// same or a nearby position) most of the 5408 haplotypes are indistinguishable: the columns of
// a 128-row batch of the config-2 matrix fall into a median of ~150 classes of bit-identical
// columns (mean ~250; ~2000 in the hypervariable regions).  With cmap_b[j] the class of column
// j in batch b and V_b[i][c] = L_i,rep(c) the value of class c in row i, an EM iteration
// (em.py:80-89 in the linear form of em.cu)
//
//     s_i = sum_j L_ij pi_j           T_j = sum_i (w_i / s_i) L_ij
//
// is evaluated exactly as
//
//     Pi_b[c] = sum_{j in class c} pi_j        (tile_pi_kernel, H adds per batch)
//     s_i     = sum_c V_b[i][c] Pi_b[c]        (tile_pass_kernel, C_b instead of H terms per row)
//     U_b[c]  = sum_{i in b} (w_i / s_i) V_b[i][c]
//     T_j     = sum_b U_b[cmap_b[j]]           (tile_gather_kernel, H adds per batch)
//
// Every sum adds non-negative terms, so regrouping them costs rounding in the last bits only
// (no cancellation); the trajectory equals that of the fp64-row pass to ~1e-13 with identical
// iteration counts (tests/test_em_gpu.py).  The pass reads N x C_b instead of N x H cells:
// 0.3 GB instead of 6.0 GB per iteration at config 2.  All orders of summation are fixed:
// results are run-to-run deterministic and identical on every rank.
//
// Classes are found from the matrix alone (the drop-in run_em sees no positions): a 64-bit
// hash of every column over the batch's rows (tile_hash_kernel), a block radix sort of
// (hash, column) per batch (tile_class_kernel; stable, so a class is a run of equal hashes
// with ascending columns), and an exact bitwise check of every cell against its class
// representative while the tiles are written (tile_fill_kernel); a batch that fails the check
// (a hash collision) sends the session back to the fp64 rows.
#pragma once

#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

#include "em_kernels.cuh"

namespace mxb {

constexpr int kTileRows = 128;       // signature rows per batch
constexpr int kTileThreads = 512;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileMaxNK = 8;        // double2 per thread and row
constexpr int kTileMaxCols = 8192;   // widest matrix (the fast path's limit on ld)
constexpr int kTileItems = 4;        // work items (row ranges) per batch, at most
constexpr int kTileClasses = 5;      // team widths 1, 2, 4, 8, 16

struct TileDesc {           // one batch
    int32_t row0, n_rows;
    int32_t n_cls, c_pad;   // classes; row stride of the batch's tile in doubles (= 64 tw nk)
    int32_t tw, nk;         // warps per row team (1, 2, 4, 8, 16), double2 per thread
    int32_t ng, pad;        // work items (contiguous row ranges) of the batch
    int64_t v_off;          // tile offset in V (doubles)
    int64_t p_off;          // offset of the batch's class sums Pi_b (doubles)
    int64_t u_off;          // offset of the batch's ng partial class vectors U_b (doubles)
};

struct TileItem {           // one work item of the pass: rows [r0, r1) of a batch
    int32_t batch, r0, r1, g;
};

__device__ __forceinline__ uint64_t tile_mix(uint64_t h, uint64_t v) {
    h = (h ^ v) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 29;
    h *= 0xC4CEB9FE1A85EC53ull;
    return h ^ (h >> 32);
}

// hash[b][j] over the rows of batch b, column j of M (row stride n_cols).
__global__ void __launch_bounds__(kTileThreads)
tile_hash_kernel(const double *__restrict__ m, int64_t n_rows, int64_t n_cols, int n_batches,
                 unsigned long long *__restrict__ hash, int hs) {
    for (int b = blockIdx.x; b < n_batches; b += gridDim.x) {
        const int64_t r0 = (int64_t)b * kTileRows;
        const int nr = (int)min((int64_t)kTileRows, n_rows - r0);
        const double *base = m + r0 * n_cols;
        for (int64_t j = threadIdx.x; j < n_cols; j += kTileThreads) {
            uint64_t h = 0x243F6A8885A308D3ull;
            const double *p = base + j;
            int r = 0;
            for (; r + 4 <= nr; r += 4) {
                const double v0 = p[(int64_t)r * n_cols], v1 = p[(int64_t)(r + 1) * n_cols];
                const double v2 = p[(int64_t)(r + 2) * n_cols], v3 = p[(int64_t)(r + 3) * n_cols];
                h = tile_mix(h, (uint64_t)__double_as_longlong(v0));
                h = tile_mix(h, (uint64_t)__double_as_longlong(v1));
                h = tile_mix(h, (uint64_t)__double_as_longlong(v2));
                h = tile_mix(h, (uint64_t)__double_as_longlong(v3));
            }
            for (; r < nr; ++r) h = tile_mix(h, (uint64_t)__double_as_longlong(p[(int64_t)r * n_cols]));
            if (h == ~0ull) h = ~0ull - 1;   // ~0 pads the sort
            hash[(size_t)b * hs + j] = h;
        }
    }
}

// Classes of one batch: sort (hash, column), number the runs of equal hashes.
//   perm[b][e]  column at sorted position e (classes are contiguous, columns ascending inside)
//   cword[b][t] run structure of entries [ITEMS t, ITEMS (t + 1)) for the class-sum kernel
//   cmap[b][j]  class of column j
//   rep[b][c]   first (smallest) column of class c
template <int ITEMS>
__global__ void __launch_bounds__(kTileThreads)
tile_class_kernel(const unsigned long long *__restrict__ hash, int hs, int n_cols, int n_batches,
                  unsigned short *__restrict__ perm, unsigned int *__restrict__ cword,
                  unsigned short *__restrict__ cmap, unsigned short *__restrict__ rep,
                  int *__restrict__ n_cls) {
    using Sort = cub::BlockRadixSort<unsigned long long, kTileThreads, ITEMS, unsigned short>;
    using Scan = cub::BlockScan<int, kTileThreads>;
    extern __shared__ __align__(128) unsigned char tile_smem[];
    typename Sort::TempStorage &sort_tmp = *reinterpret_cast<typename Sort::TempStorage *>(tile_smem);
    typename Scan::TempStorage &scan_tmp = *reinterpret_cast<typename Scan::TempStorage *>(tile_smem);
    __shared__ unsigned long long last_key[kTileThreads];
    __shared__ int first_start[kTileThreads + 1];
    const int tid = threadIdx.x;
    for (int b = blockIdx.x; b < n_batches; b += gridDim.x) {
        unsigned long long keys[ITEMS];
        unsigned short vals[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int j = tid * ITEMS + i;
            keys[i] = j < n_cols ? hash[(size_t)b * hs + j] : ~0ull;
            vals[i] = (unsigned short)j;
        }
        Sort(sort_tmp).Sort(keys, vals);
        __syncthreads();
        last_key[tid] = keys[ITEMS - 1];
        __syncthreads();
        int heads = 0;
        bool head[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int e = tid * ITEMS + i;
            const unsigned long long prev = i ? keys[i - 1] : (tid ? last_key[tid - 1] : ~keys[0]);
            head[i] = e < n_cols && keys[i] != prev;   // padding (~0) sorts behind every column
            heads += head[i] ? 1 : 0;
        }
        int before = 0, total = 0;
        Scan(scan_tmp).ExclusiveSum(heads, before, total);
        int cls = before - 1;
        unsigned int starts = 0;             // bit i: a class begins at entry i of this chunk
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int e = tid * ITEMS + i;
            if (e < n_cols) {
                if (head[i]) {
                    ++cls;
                    rep[(size_t)b * hs + cls] = vals[i];
                    starts |= 1u << i;
                }
                perm[(size_t)b * hs + e] = vals[i];
                cmap[(size_t)b * hs + vals[i]] = (unsigned short)cls;
            } else {
                starts |= 1u << i;           // past the last column: every run has ended
            }
        }
        // chunk word of the class-sum kernel: bits 0..ITEMS = run starts (bit ITEMS: at the first
        // entry of the next chunk), bits 17.. = class of the chunk's first entry
        first_start[tid] = (int)(starts & 1u);
        if (tid == 0) first_start[kTileThreads] = 1;
        __syncthreads();
        starts |= (unsigned int)first_start[tid + 1] << ITEMS;
        const int first_cls = head[0] ? before : before - 1;
        cword[(size_t)b * kTileThreads + tid] = starts | ((unsigned int)max(first_cls, 0) << 17);
        if (tid == 0) n_cls[b] = total;
        __syncthreads();   // temp storage and last_key are reused by the next batch
    }
}

// Tiles V_b[r][c] = exp(M[row0 + r][rep_b[c]] - rowmax) (the value to_linear_kernel gives
// every member of the class), zero padded to c_pad, with the exact check that makes the
// hashing safe: every cell of M must equal the cell of its class representative bit for bit.
// One warp per row.
__global__ void __launch_bounds__(kTileThreads)
tile_fill_kernel(const double *__restrict__ m, int64_t n_cols, const TileDesc *__restrict__ desc,
                 int n_batches, const unsigned short *__restrict__ cmap,
                 const unsigned short *__restrict__ rep, int hs, double *__restrict__ v,
                 int *__restrict__ bad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = blockIdx.x; b < n_batches; b += gridDim.x) {
        const TileDesc d = desc[b];
        const unsigned short *cm = cmap + (size_t)b * hs;
        const unsigned short *rp = rep + (size_t)b * hs;
        int mismatch = 0;
        for (int r = warp; r < d.n_rows; r += kTileWarps) {
            const double *row = m + (int64_t)(d.row0 + r) * n_cols;
            double mx = -INFINITY;
            for (int c = lane; c < d.n_cls; c += 32) mx = fmax(mx, row[rp[c]]);
            mx = warp_max(mx);
            for (int64_t j = lane; j < n_cols; j += 32)
                mismatch |= __double_as_longlong(row[j]) != __double_as_longlong(row[rp[cm[j]]]);
            double *dst = v + d.v_off + (int64_t)r * d.c_pad;
            for (int c = lane; c < d.c_pad; c += 32)
                dst[c] = c < d.n_cls ? exp(row[rp[c]] - mx) : 0.0;
        }
        if (__any_sync(0xffffffffu, mismatch) && lane == 0) bad[b] = 1;
    }
}

// Pi_b[c] = sum of pi_j over the members of class c.  `perm` lists the columns of a batch class
// by class (ckey = class of every entry, non-decreasing), so a class is a run of entries and
// the sums are a segmented reduction: every thread adds the entries of its chunk run by run
// in order (branch-free: the run structure is data, not control flow), stores the runs that
// begin and end inside the chunk, and the runs that cross chunk borders are completed by a
// segmented scan over the threads (shuffles in the warp, one more shuffle scan over the 16
// warp totals).  Fixed order of additions: deterministic.
constexpr int kPiPer = 16;       // entries per thread, at most (512 x 16 = 8192 columns)

struct SegItem {
    int flag;      // 1: a new class starts at (or inside) this item
    double val;    // sum of the trailing run
};
__device__ __forceinline__ SegItem seg_combine(const SegItem &a, const SegItem &b) {
    SegItem r;
    r.flag = a.flag | b.flag;
    r.val = b.flag ? b.val : a.val + b.val;
    return r;
}
__device__ __forceinline__ SegItem seg_shfl_up(const SegItem &v, int off) {
    SegItem o;
    o.flag = __shfl_up_sync(0xffffffffu, v.flag, off);
    o.val = __shfl_up_sync(0xffffffffu, v.val, off);
    return o;
}

template <int PER>
__global__ void __launch_bounds__(kTileThreads)
tile_pi_kernel(const unsigned short *__restrict__ perm, const unsigned int *__restrict__ cword,
               int hs, int n_cols, const TileDesc *__restrict__ desc, int n_batches,
               const double *__restrict__ pi0, const double *__restrict__ pi1,
               const EmState *__restrict__ st, double *__restrict__ pi_cls,
               int *__restrict__ next_item) {
    static_assert(PER % 4 == 0 && PER <= 16, "chunks are read as 8-byte words");
    extern __shared__ __align__(128) unsigned char tile_smem[];
    double *pi = reinterpret_cast<double *>(tile_smem);     // [n_cols]: the gathers below are
    // random 8-byte reads, cheaper from shared memory than through L1 tags
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    if (blockIdx.x == 0 && threadIdx.x < kTileClasses) next_item[threadIdx.x] = 0;  // the pass's queues
    const double *__restrict__ pi_g = st->cur ? pi1 : pi0;
    __shared__ int s_flag[kTileWarps];
    __shared__ double s_val[kTileWarps];
    __shared__ double s_carry[kTileWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < n_cols; j += kTileThreads) pi[j] = pi_g[j];
    __syncthreads();
    const int e0 = min(n_cols, tid * PER), cnt = min(n_cols, e0 + PER) - e0;
    const bool live = cnt > 0;
    for (int b = blockIdx.x; b < n_batches; b += gridDim.x) {
        double *out = pi_cls + desc[b].p_off;
        // the chunk's columns (PER 16-bit entries as 8-byte words) and its run structure
        const uint2 *pm = reinterpret_cast<const uint2 *>(perm + (size_t)b * hs + (size_t)tid * PER);
        uint2 words[PER / 4];
#pragma unroll
        for (int i = 0; i < PER / 4; ++i) words[i] = live ? pm[i] : make_uint2(0u, 0u);
        const unsigned int cw = cword[(size_t)b * kTileThreads + tid];
        const unsigned int starts = cw & 0x1FFFFu;
        int cls = (int)(cw >> 17);                  // class of the current run
        const int f_head = live ? (int)(starts & 1u) : 1;
        double s = 0.0, head_sum = 0.0;
        int head_open = 1;                           // the chunk's first run has not ended yet
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const unsigned int w = (q & 2) ? words[q >> 2].y : words[q >> 2].x;
            const int col = q < cnt ? (int)((q & 1) ? (w >> 16) : (w & 0xFFFFu)) : 0;
            const double v = pi[col];
            const bool start = q > 0 && ((starts >> q) & 1u);
            const bool ends = (starts >> (q + 1)) & 1u;
            if (q < cnt) {
                cls += start ? 1 : 0;
                s = (q == 0 || start) ? v : s + v;
                if (start) head_open = 0;
                // a run that ends here and did not come in from the previous thread is final
                if (ends && (!head_open || f_head)) out[cls] = s;
                if (head_open && ends) head_sum = s;
                if (ends) head_open = 0;
            }
        }
        const int started = (starts & ((1u << PER) - 1u)) != 0;     // a run began in this chunk
        const int head_cls = (int)(cw >> 17);
        if (live && head_open) head_sum = s;          // one run fills the chunk and goes on
        // aggregate of the chunk: trailing-run sum, and whether a run began here
        SegItem agg;
        agg.flag = live ? started : 1;
        agg.val = live ? s : 0.0;
        SegItem inc = agg;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const SegItem o = seg_shfl_up(inc, off);
            if (lane >= off) inc = seg_combine(o, inc);
        }
        __syncthreads();                       // the scratch of the previous batch has been read
        if (lane == 31) { s_flag[warp] = inc.flag; s_val[warp] = inc.val; }
        __syncthreads();
        if (warp == 0) {
            SegItem wv;
            wv.flag = lane < kTileWarps ? s_flag[lane] : 1;
            wv.val = lane < kTileWarps ? s_val[lane] : 0.0;
#pragma unroll
            for (int off = 1; off < kTileWarps; off <<= 1) {
                const SegItem o = seg_shfl_up(wv, off);
                if (lane >= off) wv = seg_combine(o, wv);
            }
            // exclusive: what is open at the end of the warps before this one
            const SegItem ex = seg_shfl_up(wv, 1);
            if (lane < kTileWarps) s_carry[lane] = lane == 0 ? 0.0 : ex.val;
        }
        __syncthreads();
        SegItem exc = seg_shfl_up(inc, 1);      // open run at the end of the previous thread
        if (lane == 0) { exc.flag = 0; exc.val = s_carry[warp]; }
        else if (!exc.flag) exc.val += s_carry[warp];
        if (live && !f_head) {
            // the first run came in from the previous thread(s): it is final here if it ended
            // inside the chunk or at its very end
            const bool ended = (starts >> 1) != 0;          // some later start bit, incl. bit PER
            if (ended) out[head_cls] = exc.val + head_sum;
        }
    }
}

// ---- the pass ------------------------------------------------------------------------------
// A work item is a contiguous range of rows of one batch (a batch is cut into up to four), handled
// by a *team* of `tw` warps (tw = 1 for at most 512 classes -- the bulk --, 2, 4, 8 or 16 for the
// wider ones); the 16 / tw teams of a CTA work on different items independently.  Every CTA is
// given a width when the tiles are built (CTAs in proportion to the work of each width, em.cu);
// its teams draw items of that width, longest first, from a device-wide queue.  Thread tt of a team owns the double2 chunks tt + 32 tw k (k < nk) of the
// class vectors: Pi_b and the item's share of U_b live in registers for the whole item, the
// share is stored straight from registers, and the item of a batch that finishes last adds
// the shares in item order into U_b -- no block-level synchronisation anywhere.
// The rows of the tile stream through a team-private shared-memory ring filled by 1-D bulk
// async copies (cp.async.bulk + mbarrier complete_tx, one instruction per row, issued by the
// team's first thread); a row costs a warp nk 16-byte shared loads, 4 nk DFMA, one shuffle
// butterfly and a division (plus one named barrier per row step for tw > 1); two rows are in
// flight per team where the registers allow it (nk <= 4).
constexpr int kTileRingBytes = 192 * 1024;     // all teams of a CTA together
constexpr int kTileMaxStages = 16;             // ring slots of a team
constexpr size_t kTilePassSmem = kTileRingBytes + kTileWarps * kTileMaxStages * sizeof(uint64_t) +
                                 4 * kTileWarps * sizeof(double) + 2 * kTileWarps * sizeof(int) + 128;

struct TilePlan {          // one per CTA
    int32_t tw;            // team width of this CTA
    int32_t klass;         // log2(tw): its teams take the work items of this width class
};

__device__ __forceinline__ void team_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int NKMAX, int RU>
__device__ __forceinline__ void tile_batch(const TileDesc &d, const TileItem &it, int tw, int team, int tt,
                                           const double *__restrict__ v,
                                           const double *__restrict__ pi_cls,
                                           const double *__restrict__ w, double *__restrict__ u_out,
                                           double *__restrict__ u_sum, int *__restrict__ done,
                                           uint32_t ring_u32, int ring_bytes, uint32_t bars_u32,
                                           double *red, int *ired, uint32_t &phase_bits,
                                           int &parity, int &bad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nk = d.nk;
    const int tthreads = 32 * tw;
    const uint32_t row_bytes = (uint32_t)d.c_pad * 8u;
    const int ns = min(kTileMaxStages, ring_bytes / (int)row_bytes);
    const bool producer = tt == 0;
    const int n_rows = it.r1 - it.r0;
    const double *tile = v + d.v_off + (size_t)it.r0 * d.c_pad;
    if (producer) {
        for (int q = 0; q < ns && q < n_rows; ++q) {
            mbar_expect_tx_u32(bars_u32 + 8u * q, row_bytes);
            bulk_load_u32(ring_u32 + (uint32_t)q * row_bytes, tile + (size_t)q * d.c_pad, row_bytes,
                          bars_u32 + 8u * q);
        }
    }
    const double2 *pcls = reinterpret_cast<const double2 *>(pi_cls + d.p_off) + tt;
    double2 p[NKMAX], u[NKMAX];
#pragma unroll
    for (int k = 0; k < NKMAX; ++k) {
        p[k] = k < nk ? pcls[k * tthreads] : make_double2(0.0, 0.0);
        u[k] = make_double2(0.0, 0.0);
    }
    // an item holds at most 32 rows (128-row batches, up to four items of at least 16 rows):
    // lane l keeps the weight of row l, so the row loop has no global load of its own
    const double w_lane = lane < n_rows ? w[d.row0 + it.r0 + lane] : 0.0;
    int q = 0;
    for (int r = 0; r < n_rows; r += RU) {
        // RU rows per step: their dependency chains (shared loads -> dot product -> butterfly ->
        // division -> column-sum update) overlap; the sums are formed in row order as before
        double2 x[RU][NKMAX];
        double dot[RU], wr[RU];
        int qs[RU];
        bool have[RU];
#pragma unroll
        for (int g = 0; g < RU; ++g) {
            have[g] = r + g < n_rows;
            wr[g] = have[g] ? __shfl_sync(0xffffffffu, w_lane, (r + g) & 31) : 0.0;
            qs[g] = q;
            if (have[g]) {
                mbar_wait_u32(bars_u32 + 8u * q, (phase_bits >> q) & 1u);
                phase_bits ^= 1u << q;
                if (++q == ns) q = 0;
            }
        }
#pragma unroll
        for (int g = 0; g < RU; ++g) {
            const uint32_t src = ring_u32 + (uint32_t)qs[g] * row_bytes + (uint32_t)tt * 16u;
            double dx = 0.0, dy = 0.0;
#pragma unroll
            for (int k = 0; k < NKMAX; ++k) {
                x[g][k] = (have[g] && k < nk) ? lds_v2_f64(src + (uint32_t)(k * tthreads) * 16u)
                                              : make_double2(0.0, 0.0);
                dx = fma(x[g][k].x, p[k].x, dx);
                dy = fma(x[g][k].y, p[k].y, dy);
            }
            dot[g] = dx + dy;
        }
#pragma unroll
        for (int g = 0; g < RU; ++g) dot[g] = warp_sum(dot[g]);
        if (tw > 1) {
            // warp totals -> red[parity][g][warp]; after the team barrier every thread adds the
            // tw totals of its team in warp order (lanes < tw fetch, butterfly, broadcast)
            if (lane == 0) {
#pragma unroll
                for (int g = 0; g < RU; ++g) red[(parity * RU + g) * kTileWarps + warp] = dot[g];
            }
            team_barrier(1 + team, tthreads);
#pragma unroll
            for (int g = 0; g < RU; ++g) {
                double t = lane < tw ? red[(parity * RU + g) * kTileWarps + team * tw + lane] : 0.0;
                for (int off = tw >> 1; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                dot[g] = __shfl_sync(0xffffffffu, t, 0);
            }
            parity ^= 1;
        }
        // every thread of the team holds its cells of these rows in registers: refill the slots
        // (tw > 1: the team barrier above orders the reads before the refill; one warp: its
        // lanes have converged in the butterfly, the explicit __syncwarp says so)
        if (tw == 1) __syncwarp();
        if (producer) {
#pragma unroll
            for (int g = 0; g < RU; ++g) {
                if (have[g] && r + g + ns < n_rows) {
                    mbar_expect_tx_u32(bars_u32 + 8u * qs[g], row_bytes);
                    bulk_load_u32(ring_u32 + (uint32_t)qs[g] * row_bytes,
                                  tile + (size_t)(r + g + ns) * d.c_pad, row_bytes,
                                  bars_u32 + 8u * qs[g]);
                }
            }
        }
#pragma unroll
        for (int g = 0; g < RU; ++g) {
            double coef = 0.0;
            if (wr[g] != 0.0) {
                coef = wr[g] / dot[g];
                bad |= (dot[g] == 0.0);
            }
#pragma unroll
            for (int k = 0; k < NKMAX; ++k) {
                u[k].x = fma(coef, x[g][k].x, u[k].x);
                u[k].y = fma(coef, x[g][k].y, u[k].y);
            }
        }
    }
    if (d.ng == 1) {        // the item is the whole batch
        double2 *uo = reinterpret_cast<double2 *>(u_sum + d.p_off) + tt;
#pragma unroll
        for (int k = 0; k < NKMAX; ++k)
            if (k < nk) uo[k * tthreads] = u[k];
        return;
    }
    // Store this item's share; the item of the batch that finishes last adds the ng shares in
    // item order (so the sum does not depend on which one that is) into U_b.  `done` counts
    // the finished items of the batch and is put back to zero by the last one.
    double2 *uo = reinterpret_cast<double2 *>(u_out + d.u_off + (size_t)it.g * d.c_pad) + tt;
#pragma unroll
    for (int k = 0; k < NKMAX; ++k)
        if (k < nk) uo[k * tthreads] = u[k];
    __threadfence();
    int last = 0;
    if (tw > 1) {
        team_barrier(1 + team, tthreads);              // every warp's stores are fenced
        // (a slot of its own: the team's next queue index goes through ired[team] with no
        // barrier between this read and that write)
        if (tt == 0) ired[kTileWarps + team] = atomicAdd(done + it.batch, 1) == d.ng - 1;
        team_barrier(1 + team, tthreads);
        last = ired[kTileWarps + team];
    } else {
        if (lane == 0) last = atomicAdd(done + it.batch, 1) == d.ng - 1;
        last = __shfl_sync(0xffffffffu, last, 0);
    }
    if (!last) return;
    __threadfence();
    const double2 *sh = reinterpret_cast<const double2 *>(u_out + d.u_off) + tt;
    double2 *us = reinterpret_cast<double2 *>(u_sum + d.p_off) + tt;
    const int row_d2 = d.c_pad >> 1;
#pragma unroll
    for (int k = 0; k < NKMAX; ++k) {
        if (k < nk) {
            double2 acc = __ldcg(sh + k * tthreads);
            for (int g = 1; g < d.ng; ++g) {
                const double2 o = __ldcg(sh + (size_t)g * row_d2 + k * tthreads);
                acc.x += o.x;
                acc.y += o.y;
            }
            us[k * tthreads] = acc;
        }
    }
    if (tt == 0) done[it.batch] = 0;
}

__global__ void __launch_bounds__(kTileThreads, 1)
tile_pass_kernel(const TileDesc *__restrict__ desc, const TileItem *__restrict__ items,
                 const TilePlan *__restrict__ plan, const int *__restrict__ class_ptr,
                 const int *__restrict__ class_items, int *__restrict__ next_item,
                 const double *__restrict__ v, const double *__restrict__ pi_cls,
                 const double *__restrict__ w, EmState *__restrict__ st,
                 double *__restrict__ u_out, double *__restrict__ u_sum, int *__restrict__ done) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(tile_smem + kTileRingBytes);   // [16][16]
    double *red = reinterpret_cast<double *>(bars + kTileWarps * kTileMaxStages);  // [2][2][16]
    int *ired = reinterpret_cast<int *>(red + 4 * kTileWarps);                     // [2][16]
    const int tid = threadIdx.x, warp = tid >> 5;
    const TilePlan pl = plan[blockIdx.x];
    const int tw = pl.tw;
    const int team = warp / tw, tt = tid - team * 32 * tw;
    const int ring_bytes = kTileRingBytes / (kTileWarps / tw);        // this team's share
    const uint32_t ring_u32 = smem_u32(tile_smem) + (uint32_t)(team * ring_bytes);
    const uint32_t bars_u32 = smem_u32(bars) + (uint32_t)(team * kTileMaxStages) * 8u;
    if (tt == 0) {
        for (int q = 0; q < kTileMaxStages; ++q) mbar_init(bars + team * kTileMaxStages + q, 1);
        mbar_init_fence();
    }
    __syncthreads();
    pdl_wait();               // the class sums of this iteration are complete
    pdl_launch_dependents();
    if (st->done) return;
    int bad = 0, parity = 0;
    uint32_t phase_bits = 0;
    // The teams of all CTAs of a width class draw their items from one list (longest first)
    // through a counter that tile_pi_kernel put back to zero: which team handles an item does
    // not change any result, so this costs no determinism and evens out the load.
    const int k = pl.klass;
    const int list0 = class_ptr[k], list1 = class_ptr[k + 1];
    const int lane = tid & 31;
    for (;;) {
        int i = 0;
        if (tw > 1) {
            if (tt == 0) ired[team] = atomicAdd(next_item + k, 1);
            team_barrier(1 + team, 32 * tw);
            i = ired[team];
            team_barrier(1 + team, 32 * tw);
        } else {
            if (lane == 0) i = atomicAdd(next_item + k, 1);
            i = __shfl_sync(0xffffffffu, i, 0);
        }
        if (list0 + i >= list1) break;
        const TileItem it = items[class_items[list0 + i]];
        const TileDesc d = desc[it.batch];
        if (d.nk <= 2)
            tile_batch<2, 2>(d, it, tw, team, tt, v, pi_cls, w, u_out, u_sum, done, ring_u32,
                          ring_bytes, bars_u32, red, ired, phase_bits, parity, bad);
        else if (d.nk <= 4)
            tile_batch<4, 2>(d, it, tw, team, tt, v, pi_cls, w, u_out, u_sum, done, ring_u32,
                          ring_bytes, bars_u32, red, ired, phase_bits, parity, bad);
        else
            tile_batch<8, 1>(d, it, tw, team, tt, v, pi_cls, w, u_out, u_sum, done, ring_u32,
                          ring_bytes, bars_u32, red, ired, phase_bits, parity, bad);
        // the next batch may prime the ring at once: the team barrier of the last row (tw > 1)
        // or the warp's own program order (tw == 1) came after every read of the ring
    }
    if (__any_sync(0xffffffffu, bad) && (tid & 31) == 0) atomicAdd(&st->bad, 1);
}

// T_j = sum over batches of U_b[cmap_b[j]], batches in ascending order; the batches are split
// into gridDim.y contiguous ranges whose partial sums the tail kernel adds in range order (its
// `partials` input).  One thread per column, eight batches in flight per thread.
constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll = 8;

__global__ void __launch_bounds__(kGatherThreads)
tile_gather_kernel(const unsigned short *__restrict__ cmap, int hs, int n_cols, int64_t ld,
                   const int64_t *__restrict__ p_off, int n_batches,
                   const double *__restrict__ u_sum, const EmState *__restrict__ st,
                   double *__restrict__ partials) {
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    const int j = blockIdx.x * kGatherThreads + threadIdx.x;
    if (j >= ld) return;
    const int part = blockIdx.y, n_part = gridDim.y;
    const int b0 = (int)((int64_t)n_batches * part / n_part);
    const int b1 = (int)((int64_t)n_batches * (part + 1) / n_part);
    double t = 0.0;
    if (j < n_cols) {
        int b = b0;
        for (; b + kGatherUnroll <= b1; b += kGatherUnroll) {
            double x[kGatherUnroll];
#pragma unroll
            for (int q = 0; q < kGatherUnroll; ++q)
                x[q] = u_sum[p_off[b + q] + cmap[(size_t)(b + q) * hs + j]];
#pragma unroll
            for (int q = 0; q < kGatherUnroll; ++q) t += x[q];
        }
        for (; b < b1; ++b) t += u_sum[p_off[b] + cmap[(size_t)b * hs + j]];
    }
    partials[(size_t)part * ld + j] = t;
}

}  // namespace mxb
