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
constexpr int kTileUsDoubles = 8192; // staging of the per-team class sums: 16 teams x 512 ... 1 x 8192
constexpr int kTileMaxRU = 4;        // rows in flight per team
constexpr size_t kTileSmemBytes = (kTileUsDoubles + 2 * kTileMaxRU * kTileWarps) * sizeof(double);

struct TileDesc {
    int32_t row0, n_rows;
    int32_t n_cls, c_pad;   // classes; row stride of the batch's tile in doubles (= 64 tw nk)
    int32_t tw, nk;         // warps per row team (1, 2, 4, 8, 16), double2 per thread
    int64_t v_off;          // tile offset in V (doubles)
    int64_t p_off;          // offset of the batch's class vectors in Pi / U (doubles)
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
//   cmap[b][j]  class of column j
//   rep[b][c]   first (smallest) column of class c
template <int ITEMS>
__global__ void __launch_bounds__(kTileThreads)
tile_class_kernel(const unsigned long long *__restrict__ hash, int hs, int n_cols, int n_batches,
                  unsigned short *__restrict__ perm, unsigned short *__restrict__ cmap,
                  unsigned short *__restrict__ rep, int *__restrict__ n_cls) {
    using Sort = cub::BlockRadixSort<unsigned long long, kTileThreads, ITEMS, unsigned short>;
    using Scan = cub::BlockScan<int, kTileThreads>;
    extern __shared__ __align__(16) unsigned char tile_smem[];
    typename Sort::TempStorage &sort_tmp = *reinterpret_cast<typename Sort::TempStorage *>(tile_smem);
    typename Scan::TempStorage &scan_tmp = *reinterpret_cast<typename Scan::TempStorage *>(tile_smem);
    __shared__ unsigned long long last_key[kTileThreads];
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
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int e = tid * ITEMS + i;
            if (e < n_cols) {
                if (head[i]) {
                    ++cls;
                    rep[(size_t)b * hs + cls] = vals[i];
                }
                perm[(size_t)b * hs + e] = vals[i];
                cmap[(size_t)b * hs + vals[i]] = (unsigned short)cls;
            }
        }
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

// Pi_b[c] = sum of pi_j over the members of class c, members added in ascending column order
// within a thread's chunk of `perm` and chunk totals combined by a segmented scan: fixed order.
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

__global__ void __launch_bounds__(kTileThreads)
tile_pi_kernel(const unsigned short *__restrict__ perm, const unsigned short *__restrict__ cmap,
               int hs, int n_cols, const TileDesc *__restrict__ desc, int n_batches,
               const double *__restrict__ pi0, const double *__restrict__ pi1,
               const EmState *__restrict__ st, double *__restrict__ pi_cls) {
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    const double *__restrict__ pi = st->cur ? pi1 : pi0;
    __shared__ int s_key[kTileThreads + 1];      // key of a thread's first element
    __shared__ int s_flag[kTileWarps];
    __shared__ double s_val[kTileWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n_cols + kTileThreads - 1) / kTileThreads;
    for (int b = blockIdx.x; b < n_batches; b += gridDim.x) {
        const unsigned short *pm = perm + (size_t)b * hs;
        const unsigned short *cm = cmap + (size_t)b * hs;
        double *out = pi_cls + desc[b].p_off;
        const int e0 = min(n_cols, tid * per), e1 = min(n_cols, e0 + per);
        // head run (may continue the previous thread's), interior runs (complete: stored
        // straight away), tail run (may continue into the next thread)
        int head_key = -1, tail_key = -1, n_runs = 0;
        double head_sum = 0.0, run_sum = 0.0;
        for (int eb = e0; eb < e1; eb += 8) {
            // eight elements at a time: the three dependent loads of each (position -> column
            // -> class and proportion) are issued side by side
            int key[8];
            double pv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int col = eb + q < e1 ? pm[eb + q] : 0;
                key[q] = eb + q < e1 ? cm[col] : -3;
                pv[q] = pi[col];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (key[q] == -3) break;
                if (key[q] != tail_key) {
                    if (n_runs == 1) head_sum = run_sum;
                    else if (n_runs > 1) out[tail_key] = run_sum;
                    if (n_runs == 0) head_key = key[q];
                    tail_key = key[q];
                    run_sum = pv[q];
                    ++n_runs;
                } else {
                    run_sum += pv[q];
                }
            }
        }
        const bool live = n_runs > 0;
        const bool single = n_runs == 1;
        if (single) head_sum = run_sum;
        __syncthreads();                       // s_key of the previous batch is no longer read
        s_key[tid] = live ? head_key : -2;
        if (tid == 0) s_key[kTileThreads] = -2;
        __syncthreads();
        // the last element before this chunk belongs to the previous live thread; chunks are
        // contiguous, so that is thread tid - 1 whenever this thread is live
        int prev_tail = -1;
        if (live && tid > 0) prev_tail = cm[pm[e0 - 1]];
        const int f_head = live ? (head_key != prev_tail) : 1;
        // aggregate of this thread's items [head][tail]
        SegItem agg;
        agg.flag = f_head | (single ? 0 : 1);
        agg.val = single ? head_sum : run_sum;
        if (!live) { agg.flag = 1; agg.val = 0.0; }
        // inclusive segmented scan over the threads (Kogge-Stone in the warp, then over warps)
        SegItem inc = agg;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            SegItem o;
            o.flag = __shfl_up_sync(0xffffffffu, inc.flag, off);
            o.val = __shfl_up_sync(0xffffffffu, inc.val, off);
            if (lane >= off) inc = seg_combine(o, inc);
        }
        if (lane == 31) { s_flag[warp] = inc.flag; s_val[warp] = inc.val; }
        __syncthreads();
        SegItem carry;                          // everything before this warp
        carry.flag = 1; carry.val = 0.0;
        for (int w = 0; w < warp; ++w) {
            SegItem o;
            o.flag = s_flag[w]; o.val = s_val[w];
            carry = seg_combine(carry, o);
        }
        // exclusive value for this thread = carry (+) inclusive of lane - 1
        SegItem exc;
        exc.flag = __shfl_up_sync(0xffffffffu, inc.flag, 1);
        exc.val = __shfl_up_sync(0xffffffffu, inc.val, 1);
        if (lane == 0) exc = carry; else exc = seg_combine(carry, exc);
        if (live) {
            const double head_total = f_head ? head_sum : exc.val + head_sum;
            const int next_key = s_key[tid + 1];          // -2: nothing follows
            if (single) {
                if (next_key != head_key) out[head_key] = head_total;
            } else {
                out[head_key] = head_total;               // the head run ended inside this chunk
                if (next_key != tail_key) out[tail_key] = run_sum;
            }
        }
    }
}

// ---- the pass ------------------------------------------------------------------------------
// One batch at a time per CTA (batches in descending order of work, dealt round robin).  A row
// is handled by a team of `tw` warps: thread tt of the team owns the double2 chunks tt + 32 tw k
// (k < nk) of the class vectors -- Pi_b and its share of U_b live in registers for the whole
// batch -- and the teams of a CTA take the rows of the batch in turn.  tw = 1 (at most 512
// classes, the bulk of the batches) needs no block-level synchronisation per row at all: a row
// costs a warp nk 16-byte loads, 4 nk DFMA, one shuffle butterfly and a division.  Wider teams
// add one named barrier per row.  RU rows are in flight per team.
__device__ __forceinline__ void team_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int NKMAX, int RU>
__device__ __forceinline__ void tile_batch(const TileDesc &d, const double *__restrict__ v,
                                           const double *__restrict__ pi_cls,
                                           const double *__restrict__ w, double *__restrict__ u_out,
                                           double *us, double *red, int &bad) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tw = d.tw, nk = d.nk;
    const int tthreads = 32 * tw;
    const int team = warp / tw, wit = warp - team * tw, n_teams = kTileWarps / tw;
    const int tt = tid - team * tthreads;
    const int c_pad = d.c_pad;
    const double2 *pcls = reinterpret_cast<const double2 *>(pi_cls + d.p_off) + tt;
    double2 p[NKMAX], u[NKMAX];
#pragma unroll
    for (int k = 0; k < NKMAX; ++k) {
        p[k] = k < nk ? pcls[k * tthreads] : make_double2(0.0, 0.0);
        u[k] = make_double2(0.0, 0.0);
    }
    const double *wb = w + d.row0;
    const double2 *vt = reinterpret_cast<const double2 *>(v + d.v_off) + tt;
    const int row_d2 = c_pad >> 1;
    int parity = 0;
    for (int r0 = team; r0 < d.n_rows; r0 += n_teams * RU) {
        double2 x[RU][NKMAX];
        double dot[RU], wr[RU];
#pragma unroll
        for (int g = 0; g < RU; ++g) {
            const int r = r0 + g * n_teams;
            const bool have = r < d.n_rows;
            wr[g] = have ? wb[r] : 0.0;
#pragma unroll
            for (int k = 0; k < NKMAX; ++k)
                x[g][k] = (have && k < nk) ? __ldg(vt + (size_t)r * row_d2 + k * tthreads)
                                           : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int g = 0; g < RU; ++g) {
            double dx = 0.0, dy = 0.0;
#pragma unroll
            for (int k = 0; k < NKMAX; ++k) {
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
    // class sums of the teams, added in team order
    double2 *us2 = reinterpret_cast<double2 *>(us);
    __syncthreads();   // `us` of the previous batch has been read
#pragma unroll
    for (int k = 0; k < NKMAX; ++k)
        if (k < nk) us2[team * row_d2 + tt + k * tthreads] = u[k];
    __syncthreads();
    double2 *uo = reinterpret_cast<double2 *>(u_out + d.p_off);
    for (int c = tid; c < row_d2; c += kTileThreads) {
        double2 t = us2[c];
        for (int q = 1; q < n_teams; ++q) {
            const double2 o = us2[q * row_d2 + c];
            t.x += o.x;
            t.y += o.y;
        }
        uo[c] = t;
    }
}

__global__ void __launch_bounds__(kTileThreads)
tile_pass_kernel(const TileDesc *__restrict__ desc, const int *__restrict__ order, int n_work,
                 const double *__restrict__ v, const double *__restrict__ pi_cls,
                 const double *__restrict__ w, EmState *__restrict__ st,
                 double *__restrict__ u_out) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    double *us = reinterpret_cast<double *>(tile_smem);                  // [kTileUsDoubles]
    double *red = us + kTileUsDoubles;                                   // [2][4][16]
    pdl_wait();               // Pi of this iteration is complete
    pdl_launch_dependents();
    if (st->done) return;
    int bad = 0;
    for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const TileDesc d = desc[order[wi]];
        if (d.nk <= 2) tile_batch<2, 4>(d, v, pi_cls, w, u_out, us, red, bad);
        else if (d.nk <= 4) tile_batch<4, 2>(d, v, pi_cls, w, u_out, us, red, bad);
        else tile_batch<8, 1>(d, v, pi_cls, w, u_out, us, red, bad);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicAdd(&st->bad, 1);
}

// T_j = sum over batches of U_b[cmap_b[j]], batches in ascending order; the batches are
// split into gridDim.y contiguous ranges whose partial sums the tail kernel adds in range
// order (its `partials` input).  One thread per column.
__global__ void __launch_bounds__(256)
tile_gather_kernel(const unsigned short *__restrict__ cmap, int hs, int n_cols, int64_t ld,
                   const TileDesc *__restrict__ desc, int n_batches,
                   const double *__restrict__ u_cls, const EmState *__restrict__ st,
                   double *__restrict__ partials) {
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= ld) return;
    const int part = blockIdx.y, n_part = gridDim.y;
    const int b0 = (int)((int64_t)n_batches * part / n_part);
    const int b1 = (int)((int64_t)n_batches * (part + 1) / n_part);
    double t = 0.0;
    if (j < n_cols) {
        int b = b0;
        for (; b + 4 <= b1; b += 4) {
            double x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                x[q] = u_cls[desc[b + q].p_off + cmap[(size_t)(b + q) * hs + j]];
#pragma unroll
            for (int q = 0; q < 4; ++q) t += x[q];
        }
        for (; b < b1; ++b) t += u_cls[desc[b].p_off + cmap[(size_t)b * hs + j]];
    }
    partials[(size_t)part * ld + j] = t;
}

}  // namespace mxb
