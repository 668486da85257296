// Host execution of the EM pass kernels (csrc/em_kernels.cuh compiled with MXB_CPU_EMUL) under
// the interleaving model of cuda_emul.h, against a direct long-double evaluation of
//     s_i = sum_j L_ij pi_j,   T_j = sum_i (w_i / s_i) L_ij.
// TEST INFRASTRUCTURE ONLY.  Covers what a GPU-less box can check of the kernels: the index
// arithmetic of every layout (fp64 rows, coded records, chunk-coded records for 512- and
// 384-thread CTAs, compacted records, restart pairs), ragged last chunks, odd and even row
// counts per CTA, rings shorter than the CTA's row range, the mbarrier protocols (deadlock,
// write-after-read and read-after-write hazards on the ring slots), and the packers.
#define MXB_CPU_EMUL 1
#include "cuda_emul.h"

namespace mxb {
alignas(128) unsigned char smem_raw[emul::kSmemBytes];
unsigned short cell_slot[8192];
unsigned short chunk_slot[4096];
}  // namespace mxb

#include "em_kernels.cuh"

#include <algorithm>
#include <string>

using namespace mxb;

namespace {

struct Problem {
    int64_t n_rows, n_cols, ld;
    std::vector<double> lin;      // [n_rows][ld], padding columns 0
    std::vector<double> w;        // [n_rows]
    std::vector<double> pi_a, pi_b;   // [ld], padding 0
    std::vector<long double> t_a, t_b;    // reference column sums
};

Problem make_problem(int64_t n_rows, int64_t n_cols, unsigned seed, int n_dense_rows) {
    Problem p;
    p.n_rows = n_rows;
    p.n_cols = n_cols;
    p.ld = (n_cols + 15) / 16 * 16;
    std::mt19937 r(seed);
    std::uniform_real_distribution<double> u(1e-9, 1.0);
    p.lin.assign((size_t)(n_rows * p.ld), 0.0);
    for (int64_t i = 0; i < n_rows; ++i) {
        double *row = &p.lin[(size_t)(i * p.ld)];
        if (i < n_dense_rows * 3 && i % 3 == 1) {          // a row with all-distinct values
            for (int64_t j = 0; j < n_cols; ++j) row[j] = u(r);
        } else {                                            // few values, in runs
            const int d = 3 + (int)(r() % 120);
            std::vector<double> vals((size_t)d);
            for (double &v : vals) v = u(r);
            vals[0] = 1.0;
            int64_t j = 0;
            while (j < n_cols) {
                const int64_t len = 1 + (int64_t)(r() % 40);
                const double v = vals[r() % (unsigned)d];
                for (int64_t k = j; k < std::min(n_cols, j + len); ++k) row[k] = v;
                j += len;
            }
        }
    }
    p.w.resize((size_t)n_rows);
    for (double &x : p.w) x = (double)(r() % 9);           // some zero weights
    auto props = [&](std::vector<double> &pi) {
        pi.assign((size_t)p.ld, 0.0);
        double tot = 0.0;
        for (int64_t j = 0; j < n_cols; ++j) { pi[(size_t)j] = -log(u(r)); tot += pi[(size_t)j]; }
        for (int64_t j = 0; j < n_cols; ++j) pi[(size_t)j] /= tot;
    };
    props(p.pi_a);
    props(p.pi_b);
    auto reference = [&](const std::vector<double> &pi, std::vector<long double> &t) {
        t.assign((size_t)p.ld, 0.0L);
        for (int64_t i = 0; i < n_rows; ++i) {
            const double *row = &p.lin[(size_t)(i * p.ld)];
            long double s = 0.0L;
            for (int64_t j = 0; j < n_cols; ++j) s += (long double)row[j] * pi[(size_t)j];
            if (p.w[(size_t)i] == 0.0) continue;
            const long double c = (long double)p.w[(size_t)i] / s;
            for (int64_t j = 0; j < n_cols; ++j) t[(size_t)j] += c * row[j];
        }
    };
    reference(p.pi_a, p.t_a);
    reference(p.pi_b, p.t_b);
    return p;
}

int check(const char *what, const Problem &p, const std::vector<double> &partials, int n_part,
          const std::vector<long double> &ref) {
    double worst = 0.0;
    for (int64_t j = 0; j < p.n_cols; ++j) {
        long double t = 0.0L;
        for (int c = 0; c < n_part; ++c) t += partials[(size_t)(c * p.ld + j)];
        const double err = (double)fabsl(t - ref[(size_t)j]) / (double)std::max(1e-300L, fabsl(ref[(size_t)j]));
        worst = std::max(worst, err);
    }
    const bool ok = worst < 1e-12;
    printf("%-66s %s  (max relative error %.2e)\n", what, ok ? "ok" : "FAILED", worst);
    return ok ? 0 : 1;
}

constexpr size_t kFixed = 2 * kPassWarps * kPassGroup * sizeof(double) + 16 * sizeof(uint64_t) + 256;

struct Packed {
    std::vector<unsigned char> rec;
    std::vector<double> w_coded;
    std::vector<int> flag;
    std::vector<double> dense, w_dense;
    int64_t n_dense = 0, n_coded = 0;
    size_t rec_bytes = 0;
};

// what em_pack_rows does on the host side, with the kernels run under the model
Packed pack(const Problem &p, int pair_threads, bool compact, unsigned seed) {
    Packed k;
    const int64_t n = p.n_rows;
    k.rec_bytes = pair_threads ? (size_t)pair_rec_bytes(pair_threads) : (size_t)p.ld + kDictSize * 8;
    k.rec.assign((size_t)n * k.rec_bytes, 0xCD);
    k.w_coded.assign((size_t)n, -1.0);
    k.flag.assign((size_t)n, -1);
    const int grid = (int)std::min<int64_t>(n, 5);
    if (pair_threads)
        emul::launch(dim3(grid), dim3(kPackThreads), [&] {
            em_pack_pairs_kernel(p.lin.data(), n, p.ld, p.w.data(), k.rec.data(), pair_threads,
                                 k.flag.data(), k.w_coded.data());
        }, seed, false);
    else
        emul::launch(dim3(grid), dim3(kPackThreads), [&] {
            em_pack_kernel(p.lin.data(), n, p.ld, p.w.data(), k.rec.data(), (int64_t)k.rec_bytes,
                           k.flag.data(), k.w_coded.data());
        }, seed, false);
    std::vector<int64_t> list((size_t)n, -1);
    int64_t count = -1;
    emul::launch(dim3(1), dim3(1024), [&] {
        em_dense_list_kernel(k.flag.data(), n, list.data(), &count);
    }, seed + 1, false);
    k.n_dense = count;
    k.dense.assign((size_t)(k.n_dense * p.ld), -7.0);
    k.w_dense.assign((size_t)k.n_dense, -7.0);
    if (k.n_dense > 0)
        emul::launch(dim3(2), dim3(256), [&] {
            em_gather_rows_kernel(p.lin.data(), p.ld, p.w.data(), list.data(), k.n_dense,
                                  k.dense.data(), k.w_dense.data());
        }, seed + 2, false);
    k.n_coded = n;
    if (compact && k.n_dense > 0) {
        const int64_t keep = n - k.n_dense;
        emul::launch(dim3(2), dim3(256), [&] { em_flag_invert_kernel(k.flag.data(), n); }, seed + 3, false);
        emul::launch(dim3(1), dim3(1024), [&] {
            em_dense_list_kernel(k.flag.data(), n, list.data(), &count);
        }, seed + 4, false);
        if (count != keep) { printf("compact: coded-row count %lld != %lld\n", (long long)count, (long long)keep); exit(1); }
        std::vector<unsigned char> rec2((size_t)keep * k.rec_bytes, 0xEE);
        std::vector<double> w2((size_t)keep, -3.0);
        emul::launch(dim3(3), dim3(256), [&] {
            em_gather_records_kernel(k.rec.data(), (int64_t)k.rec_bytes, k.w_coded.data(), list.data(),
                                     keep, rec2.data(), w2.data());
        }, seed + 5, false);
        k.rec.swap(rec2);
        k.w_coded.swap(w2);
        k.n_coded = keep;
    }
    return k;
}

// One pass of an iteration the way enqueue_iteration (csrc/em.cu) issues it: the fp64 pass over
// the dense rows first, then the coded pass on top (accumulate), column sums summed over CTAs.
template <int NC>
void dense_pass(const Problem &p, const Packed &k, int grid, EmState *st, std::vector<double> &partials,
                unsigned seed, bool eager) {
    const uint32_t row_bytes = (uint32_t)(p.ld * 8);
    const int stages = (int)std::min<size_t>(3, (emul::kSmemBytes - kFixed) / row_bytes);
    emul::launch(dim3(grid), dim3(kPassThreads), [&] {
        em_pass_fast_kernel<NC>((const unsigned char *)k.dense.data(), row_bytes, p.ld, k.n_dense,
                                k.w_dense.data(), p.pi_a.data(), p.pi_b.data(), st, partials.data(),
                                stages, 0);
    }, seed, eager);
}

template <int NC, class Kernel>
int single_restart(const char *name, const Problem &p, const Packed &k, int grid, int threads,
                   int stages, unsigned seed, bool eager, Kernel kernel) {
    EmState st;
    memset(&st, 0, sizeof(st));
    std::vector<double> partials((size_t)(grid * p.ld), -99.0);
    if (k.n_dense > 0) dense_pass<NC>(p, k, grid, &st, partials, seed, eager);
    emul::launch(dim3(grid), dim3(threads), [&] {
        kernel(k.rec.data(), (uint32_t)k.rec_bytes, p.ld, k.n_coded, k.w_coded.data(), p.pi_a.data(),
               p.pi_b.data(), &st, partials.data(), stages, k.n_dense > 0 ? 1 : 0);
    }, seed + 7, eager);
    if (st.bad) { printf("%s: bad rows flagged\n", name); return 1; }
    std::string what = std::string(name) + (eager ? "  [copies land early]" : "  [copies land late]");
    return check(what.c_str(), p, partials, grid, p.t_a);
}

template <int NC5, int NC3>
int run_shape(int64_t n_rows, int64_t n_cols, int grid, int stages, unsigned seed) {
    int bad = 0;
    const Problem p = make_problem(n_rows, n_cols, seed, 2);
    printf("-- %lld rows x %lld columns (ld %lld), %d CTAs, ring of %d slots, seed %u\n",
           (long long)n_rows, (long long)n_cols, (long long)p.ld, grid, stages, seed);
    const Packed cells = pack(p, 0, false, seed);
    const Packed cells_compact = pack(p, 0, true, seed);
    const Packed pairs512 = pack(p, 512, false, seed);
    const Packed pairs384 = pack(p, 384, false, seed);
    const Packed pairs512_compact = pack(p, 512, true, seed);
    printf("   dense rows: %lld (cell dictionary), %lld (chunk dictionary)\n",
           (long long)cells.n_dense, (long long)pairs512.n_dense);
    for (int eager = 0; eager < 2; ++eager) {
        const bool e = eager != 0;
        {   // the validated kernels first: fp64 rows of the whole matrix
            EmState st;
            memset(&st, 0, sizeof(st));
            std::vector<double> partials((size_t)(grid * p.ld), -99.0);
            const uint32_t row_bytes = (uint32_t)(p.ld * 8);
            const int fs = (int)std::min<size_t>(3, (emul::kSmemBytes - kFixed) / row_bytes);
            emul::launch(dim3(grid), dim3(kPassThreads), [&] {
                em_pass_fast_kernel<NC5>((const unsigned char *)p.lin.data(), row_bytes, p.ld, p.n_rows,
                                         p.w.data(), p.pi_a.data(), p.pi_b.data(), &st,
                                         partials.data(), fs, 0);
            }, seed, e);
            bad += check(e ? "em_pass_fast_kernel (fp64 rows)  [copies land early]"
                           : "em_pass_fast_kernel (fp64 rows)  [copies land late]", p, partials, grid, p.t_a);
        }
        bad += single_restart<NC5>("em_pass_coded_kernel, 512 threads", p, cells, grid, 512, stages, seed, e,
                                   em_pass_coded_kernel<NC5, 512>);
        bad += single_restart<NC5>("em_pass_coded_kernel, 384 threads", p, cells, grid, 384, stages, seed, e,
                                   em_pass_coded_kernel<NC3, 384>);
        bad += single_restart<NC5>("em_pass_coded_v3_kernel (pipelined rows)", p, cells, grid, 512, stages,
                                   seed, e, em_pass_coded_v3_kernel<NC5>);
        bad += single_restart<NC5>("em_pass_coded_v3_kernel, 384 threads", p, cells, grid, 384, stages, seed, e,
                                   em_pass_coded_v3_kernel<NC3, 384, false>);
        bad += single_restart<NC5>("em_pass_coded_v3_kernel over chunk records, 512 threads", p, pairs512, grid,
                                   512, stages, seed, e, em_pass_coded_v3_kernel<NC5, 512, true>);
        bad += single_restart<NC5>("em_pass_coded_v3_kernel over chunk records, 384 threads", p, pairs384, grid,
                                   384, stages, seed, e, em_pass_coded_v3_kernel<NC3, 384, true>);
        bad += single_restart<NC5>("em_pass_coded_kernel over coded rows only", p, cells_compact, grid, 512,
                                   stages, seed, e, em_pass_coded_kernel<NC5, 512>);
        bad += single_restart<NC5>("em_pass_coded_pairs_kernel, 512 threads", p, pairs512, grid, 512, stages,
                                   seed, e, em_pass_coded_pairs_kernel<NC5, 512>);
        bad += single_restart<NC5>("em_pass_coded_pairs_kernel, 384 threads", p, pairs384, grid, 384, stages,
                                   seed, e, em_pass_coded_pairs_kernel<NC3, 384>);
        bad += single_restart<NC5>("em_pass_coded_pairs_kernel over coded rows only", p, pairs512_compact,
                                   grid, 512, stages, seed, e, em_pass_coded_pairs_kernel<NC5, 512>);
        if (NC5 <= 6) {   // restart pairs (kMaxPairNC)
            constexpr int NP = NC5 <= 6 ? NC5 : 6;
            EmState st[2];
            memset(st, 0, sizeof(st));
            std::vector<double> pa((size_t)(grid * p.ld), -99.0), pb((size_t)(grid * p.ld), -99.0);
            const uint32_t row_bytes = (uint32_t)(p.ld * 8);
            const int fs = (int)std::min<size_t>(3, (emul::kSmemBytes - kFixed) / row_bytes);
            emul::launch(dim3(grid), dim3(kPassThreads), [&] {
                em_pass_pair_kernel<NP, false>(p.lin.data(), p.ld, p.n_rows, p.w.data(), p.pi_a.data(),
                                               p.pi_a.data(), p.pi_b.data(), p.pi_b.data(), st,
                                               pa.data(), pb.data(), fs);
            }, seed + 11, e);
            bad += check("em_pass_pair_kernel (fp64 rows), restart a", p, pa, grid, p.t_a);
            bad += check("em_pass_pair_kernel (fp64 rows), restart b", p, pb, grid, p.t_b);
            std::fill(pa.begin(), pa.end(), -99.0);
            std::fill(pb.begin(), pb.end(), -99.0);
            const Packed &k = pairs512;
            emul::launch(dim3(grid), dim3(kPassThreads), [&] {
                em_pass_pair_coded_kernel<NP>(k.rec.data(), p.ld, k.n_coded, k.w_coded.data(),
                                              p.pi_a.data(), p.pi_a.data(), p.pi_b.data(), p.pi_b.data(),
                                              st, pa.data(), pb.data(), stages);
            }, seed + 12, e);
            if (k.n_dense > 0)
                emul::launch(dim3(grid), dim3(kPassThreads), [&] {
                    em_pass_pair_kernel<NP, true>(k.dense.data(), p.ld, k.n_dense, k.w_dense.data(),
                                                  p.pi_a.data(), p.pi_a.data(), p.pi_b.data(),
                                                  p.pi_b.data(), st, pa.data(), pb.data(), fs);
                }, seed + 13, e);
            bad += check("em_pass_pair_coded_kernel + dense pair pass, restart a", p, pa, grid, p.t_a);
            bad += check("em_pass_pair_coded_kernel + dense pair pass, restart b", p, pb, grid, p.t_b);
            if (st[0].bad || st[1].bad) { printf("pair kernels flagged bad rows\n"); ++bad; }
        }
    }
    return bad;
}

}  // namespace

int main(int argc, char **argv) {
    const bool full = argc > 1 && std::string(argv[1]) == "full";
    int bad = 0;
    // narrow rows with padding columns and a ragged last chunk; odd and even row counts per CTA
    bad += run_shape<2, 2>(23, 1030, 3, 3, 1);
    // one row per CTA (only the peeled last row runs), a single CTA with an odd row count
    bad += run_shape<2, 2>(2, 1030, 2, 3, 4);
    bad += run_shape<2, 2>(7, 1030, 1, 3, 5);
    if (full) bad += run_shape<2, 2>(40, 1030, 3, 5, 6);
    // Build-17 width: 6 chunks per thread at 512 threads (ragged), 8 at 384
    bad += run_shape<6, 8>(full ? 41 : 17, 5408, full ? 3 : 2, 4, 2);
    if (full) bad += run_shape<4, 6>(29, 4096, 2, 16, 3);
    if (argc > 1 && std::string(argv[1]) == "sweep") {   // a longer one-off run over more schedules
        for (unsigned seed = 10; seed < 10 + (argc > 2 ? (unsigned)atoi(argv[2]) : 6u); ++seed) {
            bad += run_shape<2, 2>(9 + seed % 7, 1030, 1 + (int)(seed % 3), 3 + (int)(seed % 4), seed);
            bad += run_shape<3, 3>(8 + seed % 5, 2100, 1 + (int)(seed % 2), 3 + (int)(seed % 3), seed + 100);
            bad += run_shape<3, 4>(6 + seed % 6, 3000, 2, 3 + (int)(seed % 5), seed + 200);
        }
    }
    printf(bad ? "FAILED: %d checks\n" : "all checks passed (%d failures)\n", bad);
    printf("fiber switches: %ld\n", emul::switches);
    return bad ? 1 : 0;
}
