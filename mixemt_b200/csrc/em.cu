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

#include "common.cuh"
#include "em_kernels.cuh"
#include "em_tiles.cuh"
#include "tile_plan.h"


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
    // Class tiles (em_tiles.cuh): when `tiled` the pass reads `tile_v` instead of `lin`.
    bool tiled = false;
    int n_batches = 0, tile_hs = 0, tile_parts = 0, tile_grid = 0;
    unsigned char *tile_block = nullptr;   // [perm][cmap][cword][desc]
    unsigned char *tile_vecs = nullptr;    // [pi_cls][u_sum][ctas][segs][copies][gather entries]
    unsigned short *tile_perm = nullptr, *tile_cmap = nullptr;
    unsigned int *tile_cword = nullptr;
    TileDesc *tile_desc = nullptr;
    TileCta *tile_ctas = nullptr;           // plan of the pass (em_tiles.cuh)
    TileSeg *tile_segs = nullptr;
    int *tile_ent_batch = nullptr;          // gather entries: batch and offset of every U vector
    int64_t *tile_ent_off = nullptr;
    int tile_n_ent = 0, tile_n_slots = 0;
    uint32_t tile_slot_bytes = 0;
    double *tile_pi = nullptr, *tile_usum = nullptr;
    double *tile_v = nullptr;
    int64_t tile_cells = 0;                // doubles in tile_v
    int64_t tile_bytes_per_pass = 0;
    unsigned char *small = nullptr;  // one device block behind weights ... state (fewer driver calls)
    bool zero_iter = false;  // last iterate() ran no iteration
};

namespace mxb {

typedef void (*pass_fn)(const unsigned char *, uint32_t, int64_t, int64_t, const double *,
                        const double *, const double *, EmState *, double *, int);

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
constexpr int kMaxPairNC = 6;   // 2 restarts x (pi + T) x NC double2 must fit 128 registers
constexpr int kMaxSlots = 2;

// Launch with the programmatic-stream-serialization attribute (see pdl_wait): the kernel may
// be scheduled before its predecessor in the stream has finished.  MXB_EM_NO_PDL=1 turns the
// attribute off (plain stream order).
// (set before a call to launch that kernel in plain stream order)
static thread_local bool g_plain_launch = false;
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t s, Args... args) {
    static const bool pdl_on = getenv("MXB_EM_NO_PDL") == nullptr;
    const bool use_pdl = pdl_on && !g_plain_launch;
    g_plain_launch = false;
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
// `marks` (optional, 3 events) are recorded after the class sums, after the pass and after the
// gather of a tiled iteration (after the pass only for fp64 rows): per-kernel attribution.
static int enqueue_iteration(mxb_em *em, cudaEvent_t pass_begin = nullptr,
                             cudaEvent_t pass_end = nullptr, cudaEvent_t *marks = nullptr) {
    mxb_ctx *ctx = em->ctx;
    cudaStream_t s = ctx->stream;
    if (pass_begin) MXB_CUDA(cudaEventRecord(pass_begin, s));
    if (em->n_slots == 2) {
        const size_t ps = (size_t)em->n_part * em->ld;
        {
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
    if (em->tiled) {
        {
            const int per = (int)ceil_div(em->n_cols, kTileThreads);
            auto pi_fn = per <= 4 ? tile_pi_kernel<4> : per <= 8 ? tile_pi_kernel<8>
                         : per <= 12 ? tile_pi_kernel<12> : tile_pi_kernel<16>;
            MXB_CUDA(launch_pdl(pi_fn, dim3(std::min(em->n_batches, 2 * ctx->num_sms)),
                                dim3(kTileThreads), (size_t)em->ld * sizeof(double), s,
                                (const unsigned short *)em->tile_perm,
                                (const unsigned int *)em->tile_cword, em->tile_hs,
                                (int)em->n_cols, (const TileDesc *)em->tile_desc, em->n_batches,
                                (const double *)em->pi[0], (const double *)em->pi[1],
                                (const EmState *)em->state, em->tile_pi));
        }
        if (marks) MXB_CUDA(cudaEventRecord(marks[0], s));
        MXB_CUDA(launch_pdl(tile_pass_kernel, dim3(em->tile_grid), dim3(kTilePassThreads),
                            kTilePassSmem, s, (const TileCta *)em->tile_ctas,
                            (const TileSeg *)em->tile_segs, (const double *)em->tile_v, (const double *)em->tile_pi, em->state,
                            em->tile_usum, em->tile_slot_bytes, em->tile_n_slots));
        if (marks) MXB_CUDA(cudaEventRecord(marks[1], s));
        // The gather waits in plain stream order: launched early it would sit on the SMs the
        // first CTAs of the pass leave and keep them from nothing, but measured it costs 27 us
        // per iteration (B200, config 2), while the other three launches gain from the overlap.
        g_plain_launch = true;
        MXB_CUDA(launch_pdl(tile_gather_kernel,
                            dim3((unsigned)ceil_div(em->ld, kGatherThreads), em->tile_parts),
                            dim3(kGatherThreads), 0, s, (const unsigned short *)em->tile_cmap,
                            em->tile_hs, (int)em->n_cols, em->ld, (const int *)em->tile_ent_batch,
                            (const int64_t *)em->tile_ent_off, em->tile_n_ent,
                            (const double *)em->tile_usum, (const EmState *)em->state, em->partials));
        if (marks) MXB_CUDA(cudaEventRecord(marks[2], s));
        ctx->launches += 3;
    } else if (em->fast) {
        MXB_CUDA(launch_pdl(pick_pass(em->nc), dim3(em->grid_fast), dim3(kPassThreads),
                            em->smem_bytes, s, (const unsigned char *)em->lin,
                            (uint32_t)(em->ld * sizeof(double)), em->ld, em->n_rows, em->weights,
                            em->pi[0], em->pi[1], em->state, em->partials, em->n_stages));
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
    if (marks && !em->tiled) {
        for (int i = 0; i < 3; ++i) MXB_CUDA(cudaEventRecord(marks[i], s));
    }
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

// Wall-clock stage times of the one-call entry points: collected into g_stage_ms when
// mxb_stage_timing(1) is on, printed on stderr as well with MXB_TIMING=1.
struct StageTimer {
    bool on, print;
    cudaStream_t stream;
    std::chrono::steady_clock::time_point t;
    explicit StageTimer(cudaStream_t s)
        : on(g_stage_on || getenv("MXB_TIMING") != nullptr), print(getenv("MXB_TIMING") != nullptr),
          stream(s) {
        if (on) t = std::chrono::steady_clock::now();
    }
    void mark(const char *what, int stage) {
        if (!on) return;
        cudaStreamSynchronize(stream);
        auto now = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(now - t).count();
        g_stage_ms[stage] += ms;
        if (print) fprintf(stderr, "[mxb timing] %-28s %9.3f ms\n", what, ms);
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
    dev_free(em->ctx, em->tile_block);
    dev_free(em->ctx, em->tile_vecs);
    dev_free(em->ctx, em->tile_v);
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

// Class tiles of the session's matrix (em_tiles.cuh), built straight from M: no N x ld copy of
// L is ever allocated when this succeeds.  Falls through (MXB_OK, em->tiled == false) when the
// matrix does not compress to less than half of its fp64 rows, when memory is short, or when
// a batch fails the exact class check; the caller then sets up the fp64 rows.
// MXB_EM_NO_PACK=1 keeps the fp64 rows (cross-check).
template <int ITEMS>
static cudaError_t launch_tile_class(mxb_ctx *ctx, int grid, const unsigned long long *hash, int hs,
                                     int n_cols, int n_batches, unsigned short *perm,
                                     unsigned int *cword, unsigned short *cmap,
                                     unsigned short *rep, int *n_cls) {
    using Sort = cub::BlockRadixSort<unsigned long long, kTileThreads, ITEMS, unsigned short>;
    using Scan = cub::BlockScan<int, kTileThreads>;
    const size_t smem = std::max(sizeof(typename Sort::TempStorage), sizeof(typename Scan::TempStorage));
    cudaError_t e = cudaFuncSetAttribute((const void *)tile_class_kernel<ITEMS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tile_class_kernel<ITEMS><<<grid, kTileThreads, smem, ctx->stream>>>(hash, hs, n_cols, n_batches,
                                                                        perm, cword, cmap, rep, n_cls);
    return cudaGetLastError();
}

static int em_pack_tiles(mxb_em *em) {
    mxb_ctx *ctx = em->ctx;
    if (!em->fast || em->n_rows == 0 || em->n_cols > kTileMaxCols ||
        getenv("MXB_EM_NO_PACK"))
        return MXB_OK;
    const int64_t n = em->n_rows, h = em->n_cols;
    const int nb = (int)ceil_div(n, kTileRows);
    const int hs = (int)round_up(h, 8);
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_map = up((size_t)nb * hs * sizeof(unsigned short));
    unsigned char *tmp = nullptr, *block = nullptr;
    double *v = nullptr;
    // scratch: [hash][rep][n_cls][bad]
    const size_t b_hash = up((size_t)nb * hs * sizeof(unsigned long long));
    const size_t b_int = up((size_t)nb * sizeof(int));
    int rc = MXB_OK;
    cudaError_t e = dev_alloc(ctx, (void **)&tmp, b_hash + b_map + 2 * b_int);
    const bool verbose = getenv("MXB_TIMING") != nullptr;
    auto give_up = [&](bool hard, const char *what) {
        if (verbose)
            fprintf(stderr, "[mxb tiles] not used: %s (%s)\n", what, cudaGetErrorString(e));
        if (e != cudaSuccess) cudaGetLastError();
        dev_free(ctx, tmp);
        dev_free(ctx, block);
        dev_free(ctx, v);
        if (hard) {
            set_error("em_pack_tiles (%s): %s", what, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? MXB_ERR_NOMEM : MXB_ERR_CUDA;
        }
        return MXB_OK;
    };
    if (e != cudaSuccess) return give_up(false, "scratch");
    unsigned long long *hash = reinterpret_cast<unsigned long long *>(tmp);
    unsigned short *rep = reinterpret_cast<unsigned short *>(tmp + b_hash);
    int *d_ncls = reinterpret_cast<int *>(tmp + b_hash + b_map);
    int *d_bad = reinterpret_cast<int *>(tmp + b_hash + b_map + b_int);
    // persistent: [perm][cmap][desc]; the class vectors and the plan of the pass follow in a
    // second block once the class counts are known
    const size_t b_desc = up((size_t)nb * sizeof(TileDesc));
    const size_t b_cword = up((size_t)nb * kTileThreads * sizeof(unsigned int));
    const size_t head_bytes = 2 * b_map + b_cword + b_desc;
    unsigned char *maps = nullptr;
    e = dev_alloc(ctx, (void **)&maps, head_bytes);
    if (e != cudaSuccess) return give_up(false, "maps");
    block = maps;
    unsigned short *perm = reinterpret_cast<unsigned short *>(maps);
    unsigned short *cmap = reinterpret_cast<unsigned short *>(maps + b_map);
    unsigned int *cword = reinterpret_cast<unsigned int *>(maps + 2 * b_map);
    TileDesc *d_desc = reinterpret_cast<TileDesc *>(maps + 2 * b_map + b_cword);

    const int grid = std::min(nb, ctx->num_sms * 2);
    tile_hash_kernel<<<std::min(nb, ctx->num_sms * 4), kTileThreads, 0, ctx->stream>>>(
        em->mat->data, n, h, nb, hash, hs);
    e = cudaGetLastError();
    if (e == cudaSuccess) {
        const int items = (int)ceil_div(h, kTileThreads);
        if (items <= 4) e = launch_tile_class<4>(ctx, grid, hash, hs, (int)h, nb, perm, cword, cmap, rep, d_ncls);
        else if (items <= 8) e = launch_tile_class<8>(ctx, grid, hash, hs, (int)h, nb, perm, cword, cmap, rep, d_ncls);
        else if (items <= 12) e = launch_tile_class<12>(ctx, grid, hash, hs, (int)h, nb, perm, cword, cmap, rep, d_ncls);
        else e = launch_tile_class<16>(ctx, grid, hash, hs, (int)h, nb, perm, cword, cmap, rep, d_ncls);
    }
    ctx->launches += 2;
    std::vector<int> ncls((size_t)nb);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ncls.data(), d_ncls, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost,
                            ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return give_up(true, "classes");

    // layout of the batches, ring geometry and plan of the pass (tile_plan.h)
    std::vector<TileDesc> desc;
    int64_t v_cells = 0, p_cells = 0;
    int max_r_pad = 4;
    tile_layout(ncls.data(), nb, n, desc, v_cells, p_cells, max_r_pad);
    uint32_t slot_bytes = 0;
    int n_slots = 0;
    if (!tile_ring(max_r_pad, slot_bytes, n_slots)) return give_up(false, "rows too long for the ring");
    TilePlanOut plan;
    tile_plan(desc, n, ctx->num_sms, slot_bytes, p_cells, plan);
    const std::vector<TileCta> &ctas = plan.ctas;
    const std::vector<TileSeg> &segs = plan.segs;
    const std::vector<int> &ent_batch = plan.ent_batch;
    const std::vector<int64_t> &ent_off = plan.ent_off;
    const int64_t extra_cells = plan.extra_cells;
    const int n_copy = plan.n_copy;
    const int n_cta = (int)ctas.size(), n_seg = (int)segs.size();
    const int n_ent = (int)ent_batch.size();
    const int64_t u_cells = p_cells + extra_cells;
    const double tile_bytes = (double)v_cells * 8 + 3.0 * (double)nb * hs * 2 +
                              ((double)p_cells + (double)u_cells) * 8 * 2;
    const double fp64_bytes = (double)n * (double)em->ld * 8;
    if (verbose) {
        fprintf(stderr, "[mxb tiles] plan check: %d errors\n", tile_plan_check(desc, plan));
        int64_t c_sum = 0, c_max = 0, wide = 0;
        for (int b = 0; b < nb; ++b) {
            c_sum += desc[(size_t)b].n_cls;
            c_max = std::max<int64_t>(c_max, desc[(size_t)b].n_cls);
            wide += desc[(size_t)b].lg > 5;
        }
        {
            double mx = 0, mn = 1e30; int smx = 0, cmx = 0;
            for (const TileCta &c : ctas) {
                double by = 0;
                for (int q = 0; q < c.n_segs; ++q) by += (double)segs[(size_t)(c.seg0 + q)].n_rows * segs[(size_t)(c.seg0 + q)].r_pad * 8;
                mx = std::max(mx, by); mn = std::min(mn, by); smx = std::max(smx, c.n_segs); cmx = std::max(cmx, c.n_copies);
            }
            fprintf(stderr, "[mxb tiles] per CTA: bytes min %.0f max %.0f, segments max %d, copies max %d\n", mn, mx, smx, cmx);
        }
        fprintf(stderr, "[mxb tiles] %d batches of %d rows: classes mean %.1f max %lld, %lld wide "
                "batches, tiles %.3f GB (+ maps and vectors: %.3f GB) vs fp64 rows %.3f GB; %d CTAs, "
                "%d segments, %d copies, %d slots of %u B\n",
                nb, kTileRows, (double)c_sum / nb, (long long)c_max, (long long)wide,
                (double)v_cells * 8 / 1e9, tile_bytes / 1e9, fp64_bytes / 1e9, n_cta, n_seg, n_copy,
                n_slots, slot_bytes);
    }
    if (tile_bytes > 0.5 * fp64_bytes) return give_up(false, "not worth it");

    e = dev_alloc(ctx, (void **)&v, (size_t)v_cells * sizeof(double));
    unsigned char *vecs = nullptr;
    const size_t b_pi = up((size_t)p_cells * sizeof(double));
    const size_t b_u = up((size_t)u_cells * sizeof(double));
    const size_t b_cta = up((size_t)n_cta * sizeof(TileCta));
    const size_t b_seg = up((size_t)n_seg * sizeof(TileSeg));
    const size_t b_entb = up((size_t)n_ent * sizeof(int));
    const size_t b_ento = up((size_t)n_ent * sizeof(int64_t));
    if (e == cudaSuccess)
        e = dev_alloc(ctx, (void **)&vecs, b_pi + b_u + b_cta + b_seg + b_entb + b_ento);
    if (e != cudaSuccess) { dev_free(ctx, vecs); return give_up(false, "tiles"); }
    double *pi_cls = reinterpret_cast<double *>(vecs);
    double *u_sum = reinterpret_cast<double *>(vecs + b_pi);
    unsigned char *tail = vecs + b_pi + b_u;
    TileCta *d_ctas = reinterpret_cast<TileCta *>(tail);
    TileSeg *d_segs = reinterpret_cast<TileSeg *>(tail + b_cta);
    int *d_entb = reinterpret_cast<int *>(tail + b_cta + b_seg);
    int64_t *d_ento = reinterpret_cast<int64_t *>(tail + b_cta + b_seg + b_entb);
    e = cudaMemcpyAsync(d_desc, desc.data(), (size_t)nb * sizeof(TileDesc), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_ctas, ctas.data(), (size_t)n_cta * sizeof(TileCta), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_segs, segs.data(), (size_t)n_seg * sizeof(TileSeg), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_entb, ent_batch.data(), (size_t)n_ent * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_ento, ent_off.data(), (size_t)n_ent * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(vecs, 0, b_pi + b_u, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, (size_t)nb * sizeof(int), ctx->stream);
    if (e == cudaSuccess) {
        tile_fill_kernel<<<std::min(nb, ctx->num_sms * 4), kTileThreads, 0, ctx->stream>>>(
            em->mat->data, h, d_desc, nb, cmap, rep, hs, (const double *)em->weights, v, d_bad);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    std::vector<int> bad((size_t)nb);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(bad.data(), d_bad, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute((const void *)tile_pass_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTilePassSmem);
    for (int v = 0; v < 4 && e == cudaSuccess; ++v) {
        const void *fn = v == 0 ? (const void *)tile_pi_kernel<4> : v == 1 ? (const void *)tile_pi_kernel<8>
                         : v == 2 ? (const void *)tile_pi_kernel<12> : (const void *)tile_pi_kernel<16>;
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(kTileMaxCols * sizeof(double)));
    }
    if (e != cudaSuccess) { dev_free(ctx, vecs); return give_up(true, "fill"); }
    bool any_bad = false;
    for (int b = 0; b < nb; ++b) any_bad |= bad[(size_t)b] != 0;
    if (any_bad) { dev_free(ctx, vecs); return give_up(false, "hash collision"); }

    dev_free(ctx, tmp);
    em->tiled = true;
    em->n_batches = nb;
    em->tile_hs = hs;
    em->tile_block = maps;
    em->tile_vecs = vecs;
    em->tile_perm = perm;
    em->tile_cmap = cmap;
    em->tile_cword = cword;
    em->tile_desc = d_desc;
    em->tile_ctas = d_ctas;
    em->tile_segs = d_segs;
    em->tile_ent_batch = d_entb;
    em->tile_ent_off = d_ento;
    em->tile_n_ent = n_ent;
    em->tile_slot_bytes = slot_bytes;
    em->tile_n_slots = n_slots;
    em->tile_usum = u_sum;
    em->tile_pi = pi_cls;
    em->tile_v = v;
    em->tile_cells = v_cells;
    em->tile_grid = n_cta;
    em->tile_parts = std::max(1, std::min(32, n_ent / 8));
    em->n_part = em->tile_parts;
    // what an iteration reads and writes: the tiles, perm + cmap (Pi), cmap (gather), the class
    // vectors (Pi written and read, U written and read)
    em->tile_bytes_per_pass = v_cells * 8 + 3 * (int64_t)nb * hs * 2 + 2 * (p_cells + u_cells) * 8;
    return rc;
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
    STEP(cudaStreamSynchronize(ctx->stream));
    if (e == cudaSuccess) {
        // class tiles straight from M; only a matrix that does not compress gets fp64 rows of L
        const int rc = em_pack_tiles(em);
        if (rc != MXB_OK) {
            mxb_em_destroy(em);
            return rc;
        }
        if (em->tiled) em->n_slots = 1;   // restarts run one at a time over the tiles
    }
    if (!em->tiled) STEP(dev_alloc(ctx, (void **)&em->lin, (size_t)em->n_rows * row_bytes));
    if (e == cudaSuccess && em->n_rows > 0 && !em->tiled) {
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
    int rc = MXB_OK;
    // a sharded session starts from zeroed peer mailboxes on every rank (collective)
    if (rc == MXB_OK && em->sharded && ctx->world > 1 && em->fused_tail) rc = p2p_resync(ctx);
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
    if (bytes_per_pass && em->tiled) {
        *bytes_per_pass = em->tile_bytes_per_pass;
        if (n_dense_rows) *n_dense_rows = 0;
        return MXB_OK;
    }
    if (bytes_per_pass) *bytes_per_pass = em->n_rows * row_bytes;
    if (n_dense_rows) *n_dense_rows = -1;
    return MXB_OK;
}

// One iteration in log space on the session's current proportions (see em_logspace_*):
// called when the tail flagged done = kDoneNeedsLogStep.  Leaves the control block as a normal
// tail would (iters + 1, convergence / max_iter decision, buffers swapped).
static int em_logspace_step(mxb_em *em, const EmState &seen) {
    mxb_ctx *ctx = em->ctx;
    cudaStream_t s = ctx->stream;
    const int64_t n = em->n_rows, h = em->n_cols;
    constexpr int kRowBlocks = 64;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_lse = up((size_t)n * sizeof(double)), b_col = up((size_t)h * sizeof(double));
    const size_t b_part = up((size_t)kRowBlocks * h * sizeof(double));
    unsigned char *tmp = nullptr;
    if (dev_alloc(ctx, (void **)&tmp, b_lse + 2 * b_col + b_part) != cudaSuccess) {
        cudaGetLastError();
        set_error("EM log-space step: out of device memory");
        return MXB_ERR_NOMEM;
    }
    double *lse = (double *)tmp, *colmax = (double *)(tmp + b_lse);
    double *colsum = (double *)(tmp + b_lse + b_col), *part = (double *)(tmp + b_lse + 2 * b_col);
    const double *lnp = em->lnp[seen.cur];
    const int col_blocks = (int)ceil_div(h, 256);
    const int row_grid = (int)std::min<int64_t>(n, (int64_t)ctx->num_sms * 8);
    em_logspace_rows_kernel<<<row_grid, 256, 0, s>>>(em->mat->data, n, h, lnp, lse);
    em_logspace_cols_kernel<<<dim3(col_blocks, kRowBlocks), 256, 0, s>>>(
        em->mat->data, n, h, lnp, lse, em->weights, nullptr, part);
    em_logspace_colreduce_kernel<<<col_blocks, 256, 0, s>>>(part, kRowBlocks, h, 1, colmax);
    em_logspace_cols_kernel<<<dim3(col_blocks, kRowBlocks), 256, 0, s>>>(
        em->mat->data, n, h, lnp, lse, em->weights, colmax, part);
    em_logspace_colreduce_kernel<<<col_blocks, 256, 0, s>>>(part, kRowBlocks, h, 0, colsum);
    em_logspace_update_kernel<<<1, kUpdThreads, 0, s>>>(colmax, colsum, h, em->ld, em->lnp[0],
                                                        em->lnp[1], em->pi[0], em->pi[1], em->state);
    ctx->launches += 6;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    dev_free(ctx, tmp);
    if (e != cudaSuccess) {
        set_error("EM log-space step: %s", cudaGetErrorString(e));
        return MXB_ERR_CUDA;
    }
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
    int64_t log_steps = 0;
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
            if (em->host_state[wait_slot].done == kDoneNeedsLogStep) {
                // everything queued behind the flagged iteration was a no-op: drain, redo
                // that iteration in log space, go on from the state it leaves
                MXB_CUDA(cudaStreamSynchronize(ctx->stream));
                pending[0] = pending[1] = false;
                MXB_TRY(em_logspace_step(em, em->host_state[wait_slot]));
                ++log_steps;
                MXB_CUDA(cudaMemcpyAsync(&em->host_state[wait_slot], em->state, sizeof(EmState),
                                         cudaMemcpyDeviceToHost, ctx->stream));
                MXB_CUDA(cudaStreamSynchronize(ctx->stream));
                enqueued = em->host_state[wait_slot].iters;
                if (em->host_state[wait_slot].done) {
                    fin = em->host_state[wait_slot];
                    finished = true;
                }
            } else if (em->host_state[wait_slot].done) {
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

#ifdef MXB_TILE_TRACE
extern "C" int mxb_debug_tile_trace(unsigned long long *out, size_t n_words) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, mxb::g_tile_trace, std::min(n_words * 8, sizeof(mxb::g_tile_trace)));
}
#endif

int mxb_tile_plan(const int32_t *n_cls, int64_t n_batches, int64_t n_rows, int32_t num_sms,
                  int32_t *seg_out, int64_t seg_cap, int32_t *n_cta, int32_t *n_seg,
                  int32_t *n_slots_out, int32_t *slot_bytes_out, int32_t *errors) {
    MXB_REQUIRE(n_cls != nullptr && n_batches > 0 && n_rows > (n_batches - 1) * kTileRows &&
                n_rows <= n_batches * kTileRows && num_sms > 0, "bad arguments");
    for (int64_t b = 0; b < n_batches; ++b)
        MXB_REQUIRE(n_cls[b] >= 1 && n_cls[b] <= kTileMaxCols, "class counts must be in 1..8192");
    std::vector<TileDesc> desc;
    int64_t v_cells = 0, p_cells = 0;
    int max_r_pad = 4;
    tile_layout(n_cls, (int)n_batches, n_rows, desc, v_cells, p_cells, max_r_pad);
    uint32_t slot_bytes = 0;
    int n_slots = 0;
    const bool ring_ok = tile_ring(max_r_pad, slot_bytes, n_slots);
    if (n_slots_out) *n_slots_out = n_slots;
    if (slot_bytes_out) *slot_bytes_out = (int32_t)slot_bytes;
    MXB_REQUIRE(ring_ok, "rows too long for the ring");
    TilePlanOut plan;
    tile_plan(desc, n_rows, num_sms, slot_bytes, p_cells, plan);
    if (n_cta) *n_cta = (int32_t)plan.ctas.size();
    if (n_seg) *n_seg = (int32_t)plan.segs.size();
    if (errors) *errors = tile_plan_check(desc, plan);
    if (seg_out) {
        int64_t k = 0;
        for (size_t c = 0; c < plan.ctas.size(); ++c) {
            for (int q = 0; q < plan.ctas[c].n_segs && k < seg_cap; ++q, ++k) {
                const TileSeg &g = plan.segs[(size_t)(plan.ctas[c].seg0 + q)];
                const TileDesc &d = desc[(size_t)g.batch];
                int32_t *o = seg_out + 8 * k;
                o[0] = (int32_t)c;
                o[1] = g.batch;
                o[2] = (int32_t)((g.v_off - d.v_off) / d.r_pad);     // first row inside the batch
                o[3] = g.n_rows;
                o[4] = g.fit;
                o[5] = g.n_copies;
                o[6] = g.lg;
                o[7] = g.r_pad;
            }
        }
    }
    return MXB_OK;
}

int mxb_em_profile(mxb_em *em, int64_t n_iter, float *ms_out) {
    MXB_REQUIRE(em != nullptr && ms_out != nullptr && n_iter >= 1 && n_iter <= 10000, "bad argument");
    MXB_REQUIRE(em->n_slots == 1, "single-restart sessions only");
    mxb_ctx *ctx = em->ctx;
    MXB_CUDA(cudaSetDevice(ctx->device));
    em->zero_iter = false;
    MXB_TRY(reset_state(em, (long long)1 << 60, -1.0));
    std::vector<cudaEvent_t> evs((size_t)n_iter * 5);
    for (auto &e : evs) MXB_CUDA(cudaEventCreate(&e));
    for (int64_t i = 0; i < n_iter; ++i) {
        cudaEvent_t *e = &evs[(size_t)i * 5];
        MXB_TRY(enqueue_iteration(em, e[0], nullptr, e + 1));
        MXB_CUDA(cudaEventRecord(e[4], ctx->stream));
    }
    MXB_CUDA(cudaStreamSynchronize(ctx->stream));
    double tot[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < n_iter; ++i) {
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            MXB_CUDA(cudaEventElapsedTime(&ms, evs[(size_t)i * 5 + k], evs[(size_t)i * 5 + k + 1]));
            tot[k] += ms;
        }
    }
    for (auto &e : evs) cudaEventDestroy(e);
    // fp64 rows: the three marks coincide after the pass, so [0] holds the pass
    if (!em->tiled) { tot[1] = tot[0]; tot[0] = 0.0; }
    for (int k = 0; k < 4; ++k) ms_out[k] = (float)(tot[k] / (double)n_iter);
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
    const int W = ctx->world, me = ctx->rank;
    cudaStream_t s = ctx->stream;
    if (W <= 1 || !ctx->nccl_comm) {
        if (sub_log != 0.0) {
            const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 16);
            fold_shift_kernel<<<grid, 256, 0, s>>>(m->data, n, sub_log);
            ctx->launches++;
            MXB_CUDA(cudaGetLastError());
        }
        MXB_CUDA(cudaStreamSynchronize(s));
        return MXB_OK;
    }
    // all-to-all of row shards, local logaddexp fold in rank order (restarts are
    // dealt in contiguous blocks, so rank order is restart order), shards gathered back.
    const int64_t lo = m->n_rows * me / W, hi = m->n_rows * (me + 1) / W;
    const int64_t max_rows = ceil_div(m->n_rows, W);
    const int64_t slot = max_rows * m->n_cols;
    double *recv = nullptr;
    if (dev_alloc(ctx, (void **)&recv, (size_t)W * (size_t)slot * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        set_error("mxb_matrix_fold_ranks: no memory for the %d receive slots", W);
        return MXB_ERR_NOMEM;
    }
    int rc = nccl_alltoall_rows(ctx, m->data, recv, m->n_rows, m->n_cols, slot);
    if (rc == MXB_OK && hi > lo) {
        const int64_t cells = (hi - lo) * m->n_cols;
        const int grid = (int)std::min<int64_t>(ceil_div(cells, 256), (int64_t)ctx->num_sms * 16);
        fold_shards_kernel<<<grid, 256, 0, s>>>(m->data + lo * m->n_cols, recv, slot, cells, W, me,
                                                 sub_log);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = MXB_ERR_CUDA;
    }
    if (rc == MXB_OK) rc = nccl_allgather_rows(ctx, m->data, m->n_rows, m->n_cols);
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == MXB_OK) rc = MXB_ERR_CUDA;
    dev_free(ctx, recv);
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
// matrices folded with logaddexp in restart order (em.py:156): a restart that finishes before
// an earlier one parks the proportions its read matrix is formed from until its turn.
static int run_em_batched(mxb_em *em, const double *init_lnprops, int32_t n_multi,
                          int64_t max_iter, double tol, bool raw, mxb_matrix *mix,
                          double *props_out, int64_t *iters_out, int32_t *converged_out) {
    mxb_ctx *ctx = em->ctx;
    cudaStream_t s = ctx->stream;
    const int64_t h = em->n_cols, ld = em->ld;
    // device: [inits n_multi x h][final log-proportions n_multi x h][the proportions before
    // the last step n_multi x h]; the finals are read back in one copy at the end (no pinned
    // host buffer, no blocking copy mid-run)
    const size_t vec_bytes = (size_t)n_multi * h * sizeof(double);
    double *d_inits = nullptr;
    if (dev_alloc(ctx, (void **)&d_inits, 3 * vec_bytes) != cudaSuccess) {
        set_error("run_em: device allocation of %zu bytes failed", 3 * vec_bytes);
        cudaGetLastError();
        return MXB_ERR_NOMEM;
    }
    double *d_fin = d_inits + (size_t)n_multi * h;
    double *d_prev = d_fin + (size_t)n_multi * h;
    std::vector<double> h_fin;
    std::vector<char> done_restart;       // finished, read matrix not folded yet (or folded)
    try {
        h_fin.resize((size_t)n_multi * h);
        done_restart.assign((size_t)n_multi, 0);
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
                    // the read matrix comes from the proportions before the last step (SURVEY
                    // F5); the slot is reused, so they are kept until the restarts before this
                    // one have been folded
                    if (cudaMemcpyAsync(d_prev + (size_t)idx * h, em->lnp[st.cur] + sl * ld,
                                        h * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
                        set_error("run_em: result copy failed");
                        rc = MXB_ERR_CUDA;
                        break;
                    }
                    done_restart[(size_t)idx] = 1;
                    while (folded < n_multi && done_restart[(size_t)folded]) {
                        const bool last = folded == n_multi - 1;
                        const double sub = (last && n_multi > 1 && !raw) ? log((double)n_multi) : 0.0;
                        read_mix_kernel<<<grid_mix, kMixThreads, 0, s>>>(
                            em->mat->data, em->n_rows, em->n_cols, d_prev + (size_t)folded * h,
                            mix->data, folded == 0 ? 0 : 1, sub);
                        ctx->launches++;
                        ++folded;
                    }
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
    tm.mark("em_create (tiles or fp64 rows)", kStageSetup);
    if (rc == MXB_OK && want_mix) rc = mxb_matrix_alloc(ctx, m->n_rows, h, &mix);
    tm.mark("alloc read_mix", kStageSetup);
    // (started after the device allocations: page faults and cudaMalloc contend for the
    // process's mmap lock)
    if (rc == MXB_OK && read_mix_out)
        fault_mix.start(read_mix_out, (size_t)m->n_rows * (size_t)h * sizeof(double));
    const bool batched = rc == MXB_OK && em->n_slots == 2;
    if (batched) {
        rc = run_em_batched(em, init_lnprops, n_multi, max_iter, tol, raw, mix, props_out,
                            iters_out, converged_out);
        tm.mark("iterate (two restarts per pass)", kStageIterate);
    }
    for (int32_t i = 0; rc == MXB_OK && !batched && i < n_multi; ++i) {
        int64_t iters = 0;
        int32_t conv = 0;
        rc = mxb_em_set_lnprops(em, init_lnprops + (size_t)i * h);
        if (rc == MXB_OK) rc = mxb_em_iterate(em, max_iter, tol, &iters, &conv);
        tm.mark("iterate", kStageIterate);
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
            tm.mark("read_mix kernel", kStageReadMix);
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
            tm.mark("download read_mix", kStageD2H);
        }
        else if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            set_error("mxb_run_em: stream sync failed");
            rc = MXB_ERR_CUDA;
        }
    }
    mxb_em_destroy(em);
    if (rc == MXB_OK && read_mix_dev) *read_mix_dev = mix;
    else mxb_matrix_destroy(mix);
    tm.mark("free", kStageOther);
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
