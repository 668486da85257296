"""Pins the CPU oracle (oracle/oracle_np.py, oracle/oracle.c) to the
reference: against the golden vectors the unmodified reference produced
(oracle/make_golden.py), and against the live reference when it is mounted."""
import numpy as np
import pytest

from mixemt_b200 import synth
from mixemt_b200.preprocess import HapVarBaseMatrix, parse_signatures
from oracle import oracle_c, oracle_np, refload
from conftest import load_golden, make_args


def seeded_inits(seed, n_multi, h):
    np.random.seed(seed)
    return np.array([np.log(np.random.dirichlet([1.0] * h)) for _ in range(n_multi)])


def test_build_toy_golden(toy_phylo, golden_toy):
    haps = list("ABCDEFGHI")
    for key, mat_key in (("build_reads", "build_mat"), ("em_reads", "em_mat")):
        reads = str(golden_toy[key]).split("\n")
        for fn in (oracle_np.build_matrix_loops, oracle_np.build_matrix_fast):
            mat, _ = fn("AAAAAAAAA", toy_phylo, reads, haps)
            assert np.array_equal(mat, golden_toy[mat_key])
        t = HapVarBaseMatrix("AAAAAAAAA", toy_phylo, haps).pack()
        csr, _ = parse_signatures(reads, t)
        mat, cnt = oracle_c.build_matrix(t, csr)
        assert np.array_equal(mat, golden_toy[mat_key])
        assert np.array_equal(cnt, oracle_np.build_matrix_loops("AAAAAAAAA", toy_phylo, reads, haps)[1])


@pytest.mark.parametrize("fixture,gold", [("phylo17", "golden_build17.npz"),
                                          ("phylo17_cfg5", "golden_build17_cfg5.npz")])
def test_build17_golden(fixture, gold, request):
    phylo = request.getfixturevalue(fixture)
    g = load_golden(gold)
    reads = str(g["reads"]).split("\n")
    haps = sorted(phylo.hap_var)
    t = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr, err = parse_signatures(reads, t)
    assert err is None
    mat, cnt = oracle_c.build_matrix(t, csr)
    assert np.array_equal(mat, g["mat"])                        # bit-exact
    rows = [0, len(reads) - 2, len(reads) - 1]                  # incl. the odd-base rows
    py, py_cnt = oracle_np.build_matrix_fast(phylo.refseq, phylo, [reads[i] for i in rows], haps)
    assert np.array_equal(py, g["mat"][rows]) and np.array_equal(py_cnt, cnt[rows])


def test_em_step_golden(golden_toy):
    inf = float("inf")
    in_mat = np.array([[0.0, -inf, -inf], [-inf, 0.0, -inf], [-inf, -inf, 0.0]])
    start = np.log(np.array([0.6, 0.2, 0.2]))
    for key, wts in (("step_w111", [1, 1, 1]), ("step_w211", [2, 1, 1])):
        mix, new = oracle_np.em_step(in_mat, np.array(wts), start)
        assert np.array_equal(mix, golden_toy[key + "_mix"])
        assert np.array_equal(new, golden_toy[key + "_props"])   # exact, like em_test.py:48
        mix, new = oracle_c.em_step(in_mat, np.array(wts, dtype=float), start)
        assert np.array_equal(mix, golden_toy[key + "_mix"])
        assert np.abs(new - golden_toy[key + "_props"]).max() < 3e-16
    mat = golden_toy["em_mat"]
    for impl, tol in ((oracle_np, 0.0), (oracle_c, 1e-13)):
        mix, new = impl.em_step(mat, np.arange(1, 11).astype(float), golden_toy["step_toy_start"])
        assert np.abs(new - golden_toy["step_toy_props"]).max() <= tol
        assert np.abs(mix - golden_toy["step_toy_mix"]).max() <= tol
    mix2, new2 = oracle_np.em_step_scipy(mat, np.arange(1, 11), golden_toy["step_toy_start"],
                                         np.empty_like(mat))
    assert np.array_equal(new2, golden_toy["step_toy_props"])


@pytest.mark.parametrize("n_multi", [1, 4, 10])
def test_run_em_toy_golden(n_multi, golden_toy):
    mat = golden_toy["em_mat"]
    inits = seeded_inits(int(golden_toy["run_%d_seed" % n_multi]), n_multi, 9)
    for impl, tol in ((oracle_np, 1e-15), (oracle_c, 1e-12)):
        props, mix, iters = impl.run_em(mat, np.ones(10), inits, 1000, 1e-4)
        assert list(iters) == golden_toy["run_%d_iters" % n_multi].tolist()
        assert np.abs(props - golden_toy["run_%d_props" % n_multi]).max() <= tol
        assert np.abs(mix - golden_toy["run_%d_mix" % n_multi]).max() <= max(tol, 1e-15) * 1e3


def test_run_em_exhaust_golden(golden_toy):
    inits = np.log(np.full((1, 9), 1.0 / 9))
    for impl in (oracle_np, oracle_c):
        props, mix, iters = impl.run_em(golden_toy["em_mat"], np.arange(1, 11).astype(float),
                                        inits, 7, 1e-4)
        assert list(iters) == [7]
        assert np.abs(props - golden_toy["run_exhaust_props"]).max() < 1e-14
        assert np.abs(mix - golden_toy["run_exhaust_mix"]).max() < 1e-12


def test_run_em_build17_golden(phylo17):
    """Build-17 sub-problem (600 x 512) to convergence: the C oracle reproduces
    the reference's iteration counts, proportions and assignments."""
    g = load_golden("golden_em17.npz")
    haps = sorted(phylo17.hap_var)
    haps_b = [haps[j] for j in g["b_cols"].tolist()]
    reads = str(g["b_reads"]).split("\n")
    t = HapVarBaseMatrix(phylo17.refseq, phylo17, haps_b).pack()
    csr, _ = parse_signatures(reads, t)
    mat, _ = oracle_c.build_matrix(t, csr, want_counts=False)
    for tag, n_multi in (("b1", 1), ("b3", 3)):
        inits = seeded_inits(int(g[tag + "_seed"]), n_multi, len(haps_b))
        props, mix, iters = oracle_c.run_em(mat, g["b_weights"], inits, 5000, 1e-4)
        assert list(iters) == g[tag + "_iters"].tolist()
        assert np.abs(props - g[tag + "_props"]).max() < 1e-10
        assert np.abs(mix[:6] - g[tag + "_mix_rows"]).max() < 1e-9
        assert np.array_equal(np.argmax(mix, 1), g[tag + "_argmax"])


@pytest.mark.skipif(not refload.available(), reason="reference not mounted on this box")
def test_against_live_reference(phylo17):
    """In the build container: the oracle against the reference itself."""
    _, ref_pre, ref_em = refload.load()
    haps = sorted(phylo17.hap_var)
    mix = synth.make_mixture(phylo17, phylo17.refseq, [("H1", 0.6), ("M7", 0.4)], 300, seed=9)
    reads = mix.signatures[:12]
    want = ref_pre.build_em_matrix(phylo17.refseq, phylo17, reads, haps, make_args())
    t = HapVarBaseMatrix(phylo17.refseq, phylo17, haps).pack()
    csr, _ = parse_signatures(reads, t)
    assert np.array_equal(oracle_c.build_matrix(t, csr)[0], want)
    sub = want[:, ::11].copy()
    np.random.seed(77)
    p_ref, m_ref = ref_em.run_em(sub, mix.weights[:12], make_args(n_multi=2, max_iter=200))
    inits = seeded_inits(77, 2, sub.shape[1])
    p_np, m_np, _ = oracle_np.run_em(sub, mix.weights[:12], inits, 200, 1e-4)
    p_c, m_c, _ = oracle_c.run_em(sub, mix.weights[:12], inits, 200, 1e-4)
    assert np.array_equal(p_np, p_ref) and np.array_equal(m_np, m_ref)
    assert np.abs(p_c - p_ref).max() < 1e-13 and np.abs(m_c - m_ref).max() < 1e-10


# ---- consumers of the EM result ---------------------------
def _decode_assign(res, n):
    out = np.full(n, -2, dtype=np.int32)
    for name, rows in res.items():
        out[sorted(rows)] = -1 if name == 'unassigned' else int(name[3:]) - 1
    return out


@pytest.mark.parametrize("seed", [71, 72])
def test_consumers_golden(seed):
    g = load_golden("golden_consumers.npz")
    props, mix, wts = oracle_np.synthetic_em_result(seed)
    haps = ["hg%d" % j for j in range(mix.shape[1])]
    for min_reads in (1, 60, 400):
        got = oracle_np.find_contribs_from_reads(mix, wts, min_reads)
        assert got == g["s%d_contribs_r%d" % (seed, min_reads)].tolist()
    top = g["s%d_top" % seed].tolist()
    contribs = [["hap%d" % (k + 1), haps[j], props[j]] for k, j in enumerate(top)]
    for tag, fold, cons in (("f2", 2.0, contribs), ("f1p2", 1.2, contribs[:2]),
                            ("single", 2.0, contribs[:1])):
        res = oracle_np.assign_read_indexes(cons, (props, mix), haps, len(mix), fold)
        assert np.array_equal(_decode_assign(res, len(mix)), g["s%d_assign_%s" % (seed, tag)])
    red, new_haps = oracle_np.reduce_em_matrix(mix, haps, contribs)
    assert np.array_equal(red, g["s%d_reduced" % seed])
    assert [haps.index(x) for x in new_haps] == g["s%d_reduced_cols" % seed].tolist()


@pytest.mark.skipif(not refload.available(), reason="reference not mounted")
def test_consumers_live_reference():
    refload.load()
    from mixemt import assemble, preprocess as ref_pre
    props, mix, wts = oracle_np.synthetic_em_result(99, n=150, h=40)
    haps = ["hg%d" % j for j in range(40)]
    ns = make_args(min_reads=25)
    assert assemble._find_contribs_from_reads(mix, wts, ns) == \
        oracle_np.find_contribs_from_reads(mix, wts, 25)
    top = np.argsort(props)[::-1][:3].tolist()
    contribs = [["hap%d" % (k + 1), haps[j], props[j]] for k, j in enumerate(top)]
    ref = assemble.assign_read_indexes(contribs, (props, mix), haps, list(range(150)), 2.0)
    got = oracle_np.assign_read_indexes(contribs, (props, mix), haps, 150, 2.0)
    assert dict(ref) == got
    a, b = ref_pre.reduce_em_matrix(mix, haps, contribs), oracle_np.reduce_em_matrix(mix, haps, contribs)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]


@pytest.mark.skipif(not refload.available(), reason="reference not mounted")
def test_em_step_fuzz_against_live_reference():
    """The oracles' em_step against the reference's own (em.py:57-91, scipy logsumexp) on small
    random inputs with the edge values the domain has: -inf cells, ties among the row and
    column maxima, a row that fits no haplotype (NaN, like the reference), zero weights,
    components with zero proportion.  numpy restatement: bit-identical, NaN for NaN; C port:
    the same non-finite pattern and within 1e-12."""
    import warnings
    _, _, ref_em = refload.load()
    rs = np.random.RandomState(0)
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for it in range(200):
            n, h = rs.randint(1, 14), rs.randint(1, 10)
            mat = -rs.gamma(2.0, 4.0, size=(n, h))
            mode = it % 5
            if mode >= 1:
                mat[rs.rand(n, h) < 0.2] = -np.inf
            if mode >= 2:
                mat = np.round(mat)
            if mode == 3:
                mat[rs.randint(n)] = -np.inf
            wts = rs.randint(0 if mode >= 2 else 1, 50, size=n)
            lp = np.log(rs.dirichlet([1.0] * h))
            if mode == 4:
                lp[rs.rand(h) < 0.3] = -np.inf
            z_ref, p_ref = ref_em.em_step(mat, wts, lp, np.empty_like(mat))
            z_np, p_np = oracle_np.em_step(mat, wts, lp)
            assert np.array_equal(z_np, z_ref, equal_nan=True), it
            assert np.array_equal(p_np, p_ref, equal_nan=True), it
            z_c, p_c = oracle_c.em_step(mat, wts.astype(np.float64), lp)
            for got, want in ((z_c, z_ref), (p_c, p_ref)):
                fin = np.isfinite(want)
                assert np.array_equal(got[~fin], want[~fin], equal_nan=True), it
                assert not fin.any() or np.abs(got[fin] - want[fin]).max() < 1e-12, it
