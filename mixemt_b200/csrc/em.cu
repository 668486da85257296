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
    static const bool v3 = getenv("MXB_EM_CODED_V3") != nullptr;
    static const bool t384 = getenv("MXB_EM_CODED_T384") != nullptr;
    const int nc384 = (int)ceil_div(ld / 2, 384);   // chunks per thread of a 384-thread CTA
    if (v3) {   // pipelined rows; over cell records or chunk records, 512 or 384 threads
        if (pair_threads == 384) {
            switch (nc384) {
                case 1: return {em_pass_coded_v3_kernel<1, 384, true>, 384};
                case 2: return {em_pass_coded_v3_kernel<2, 384, true>, 384};
                case 3: return {em_pass_coded_v3_kernel<3, 384, true>, 384};
                case 4: return {em_pass_coded_v3_kernel<4, 384, true>, 384};
                case 5: return {em_pass_coded_v3_kernel<5, 384, true>, 384};
                case 6: return {em_pass_coded_v3_kernel<6, 384, true>, 384};
                case 7: return {em_pass_coded_v3_kernel<7, 384, true>, 384};
                case 8: return {em_pass_coded_v3_kernel<8, 384, true>, 384};
            }
            return {nullptr, 0};
        }
        if (pair_threads != 0) {
            switch (nc) {
                case 1: return {em_pass_coded_v3_kernel<1, kPassThreads, true>, kPassThreads};
                case 2: return {em_pass_coded_v3_kernel<2, kPassThreads, true>, kPassThreads};
                case 3: return {em_pass_coded_v3_kernel<3, kPassThreads, true>, kPassThreads};
                case 4: return {em_pass_coded_v3_kernel<4, kPassThreads, true>, kPassThreads};
                case 5: return {em_pass_coded_v3_kernel<5, kPassThreads, true>, kPassThreads};
                case 6: return {em_pass_coded_v3_kernel<6, kPassThreads, true>, kPassThreads};
                case 7: return {em_pass_coded_v3_kernel<7, kPassThreads, true>, kPassThreads};
                case 8: return {em_pass_coded_v3_kernel<8, kPassThreads, true>, kPassThreads};
            }
            return {nullptr, 0};
        }
        if (t384 && nc384 >= 6 && nc384 <= 8) {
            switch (nc384) {
                case 6: return {em_pass_coded_v3_kernel<6, 384, false>, 384};
                case 7: return {em_pass_coded_v3_kernel<7, 384, false>, 384};
                case 8: return {em_pass_coded_v3_kernel<8, 384, false>, 384};
            }
        }
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
    if (pair_threads == 384) {
        switch (nc384) {
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
    if (t384 && nc384 >= 6 && nc384 <= 8) {   // other widths keep the 512-thread kernel
        switch (nc384) {
            case 6: return {em_pass_coded_kernel<6, 384>, 384};
            case 7: return {em_pass_coded_kernel<7, 384>, 384};
            case 8: return {em_pass_coded_kernel<8, 384>, 384};
        }
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
