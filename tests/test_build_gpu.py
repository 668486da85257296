"""Kernel 1 parity: CUDA build_em_matrix (through the C-ABI) against the
reference's golden outputs and the CPU oracle.  Bar: bit-exact log-likelihoods
(same fp64 additions in the same order) and bit-exact match counts."""
import numpy as np
import pytest

import mixemt_b200
from mixemt_b200 import synth
from mixemt_b200.preprocess import (HapVarBaseMatrix, build_em_matrix, build_matrix_from_csr,
                                    build_em_matrix_device, parse_signatures)
from oracle import oracle_c, oracle_np
from conftest import make_args, load_golden

pytestmark = pytest.mark.gpu


def test_build_em_matrix_simple(toy_phylo, golden_toy):
    """reference preprocess_test.py:268-284 (4 x 9, allclose) -- here exact."""
    reads = str(golden_toy["build_reads"]).split("\n")
    in_mat = build_em_matrix("AAAAAAAAA", toy_phylo, reads, list("ABCDEFGHI"), make_args())
    r1 = ([(0.01 / 3) * (0.01 / 3)] + ([0.99 * (0.01 / 3)] * 8))
    r2 = ([0.99 * (0.01 / 3)] + ([(0.01 / 3) * (0.01 / 3)] * 8))
    r3 = ([(0.98) * (0.02 / 3)] + [(0.02 / 3) * (0.98)] + [(0.02 / 3) * (0.02 / 3)] +
          [(0.02 / 3) * (0.98)] + [0.98 * 0.98] + ([(0.02 / 3) * (0.98)] * 3) +
          [(0.02 / 3) * (0.02 / 3)])
    r4 = ([0.99 * (0.02 / 3)] + [(0.01 / 3) * (0.98)] + [(0.01 / 3) * (0.02 / 3)] +
          ([(0.01 / 3) * (0.98)] * 5) + [0.99 * (0.02 / 3)])
    res_mat = np.log(np.array([r1, r2, r3, r4]))
    assert in_mat.shape == (4, 9)
    assert in_mat.dtype == np.float64 and in_mat.flags["C_CONTIGUOUS"]
    assert np.allclose(in_mat, res_mat)
    assert np.array_equal(in_mat, golden_toy["build_mat"])


def test_toy_em_matrix_exact(toy_phylo, golden_toy):
    reads = str(golden_toy["em_reads"]).split("\n")
    mat = build_em_matrix("AAAAAAAAA", toy_phylo, reads, list("ABCDEFGHI"), make_args())
    assert np.array_equal(mat, golden_toy["em_mat"])


def test_build17_golden_bit_exact(phylo17):
    gold = load_golden("golden_build17.npz")
    reads = str(gold["reads"]).split("\n")
    haps = sorted(phylo17.hap_var)
    mat = build_em_matrix(phylo17.refseq, phylo17, reads, haps, make_args())
    assert mat.shape == gold["mat"].shape == (len(reads), 5408)
    assert np.array_equal(mat, gold["mat"])


def test_build17_cfg5_golden_bit_exact(phylo17_cfg5):
    """--unstable + --exclude_pos (doubled mutation counts) + custom
    haplotypes: BASELINE.json config 5."""
    gold = load_golden("golden_build17_cfg5.npz")
    reads = str(gold["reads"]).split("\n")
    haps = sorted(phylo17_cfg5.hap_var)
    assert "custom_hap1" in haps and "custom_hap2" in haps
    mat = build_em_matrix(phylo17_cfg5.refseq, phylo17_cfg5, reads, haps, make_args())
    assert np.array_equal(mat, gold["mat"])


def test_build17_vs_oracle_counts_and_values(phylo17):
    """2000 fragments of a 3-way mixture x 5408 haplotypes: values and match
    counts bit-exact against the C oracle; counts + mismatches = K."""
    haps = sorted(phylo17.hap_var)
    mix = synth.make_mixture(phylo17, phylo17.refseq,
                             [("H1", 0.5), ("L3e", 0.3), ("U5a1", 0.2)], 2000, seed=11)
    tables = HapVarBaseMatrix(phylo17.refseq, phylo17, haps).pack()
    csr = mix.csr(tables)
    mat, cnt, _, ms = build_matrix_from_csr(tables, csr, want_counts=True)
    o_mat, o_cnt = oracle_c.build_matrix(tables, csr)
    assert np.array_equal(cnt, o_cnt)
    assert np.array_equal(mat, o_mat)
    k = np.diff(csr.row_ptr)
    assert cnt.max() <= k.max() and np.all(cnt <= k[:, None]) and np.all(cnt >= 0)
    # the string API gives the same matrix as the CSR API
    mat2 = build_em_matrix(phylo17.refseq, phylo17, mix.signatures, haps, make_args())
    assert np.array_equal(mat, mat2)
    # true sources are the best-supported columns
    votes = np.bincount(np.argmax(mat, 1), weights=mix.weights, minlength=len(haps))
    assert ms > 0.0 and votes.sum() == mix.weights.sum()


# 6000 haplotypes: the 256-group shared-memory layout of the class kernel; 8300: more than its
# 8192 columns, i.e. the dense kernel
@pytest.mark.parametrize("n_hap,n_pos", [(1, 5), (31, 40), (33, 64), (1024, 100), (1025, 333),
                                         (2500, 700), (6000, 900), (8300, 400)])
def test_build_shapes_vs_oracle(n_hap, n_pos):
    phylo, refseq = synth.synthetic_phylo(n_hap=n_hap, n_pos=n_pos, ref_len=max(2 * n_pos, 600),
                                          markers_per_hap=min(6, n_pos), seed=n_hap)
    haps = sorted(phylo.hap_var)
    mix = synth.make_mixture(phylo, refseq, [(haps[0], 0.6), (haps[-1], 0.4)], 500,
                             frag_len=150, err=0.05, seed=3)
    tables = HapVarBaseMatrix(refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    mat, cnt, _, _ = build_matrix_from_csr(tables, csr, want_counts=True)
    o_mat, o_cnt = oracle_c.build_matrix(tables, csr)
    assert np.array_equal(mat, o_mat) and np.array_equal(cnt, o_cnt)
    if n_hap <= 33:
        py_mat, py_cnt = oracle_np.build_matrix_loops(refseq, phylo, mix.signatures[:20], haps)
        assert np.array_equal(mat[:20], py_mat) and np.array_equal(cnt[:20], py_cnt)


def test_build_long_signature_chunks():
    """K > 512 observations forces the chunked accumulation path."""
    phylo, refseq = synth.synthetic_phylo(n_hap=70, n_pos=1500, ref_len=3000, seed=5)
    haps = sorted(phylo.hap_var)
    mix = synth.make_mixture(phylo, refseq, [(haps[1], 1.0)], 40, frag_len=2900, err=0.01, seed=8)
    tables = HapVarBaseMatrix(refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    assert np.diff(csr.row_ptr).max() > 1024
    mat, cnt, _, _ = build_matrix_from_csr(tables, csr, want_counts=True)
    o_mat, o_cnt = oracle_c.build_matrix(tables, csr)
    assert np.array_equal(mat, o_mat) and np.array_equal(cnt, o_cnt)


def test_build_edge_inputs(toy_phylo):
    haps = list("ABCDEFGHI")
    ref = "AAAAAAAAA"
    assert build_em_matrix(ref, toy_phylo, [], haps, make_args()).shape == (0, 9)
    assert build_em_matrix(ref, toy_phylo, ["1:A"], [], make_args()).shape == (1, 0)
    # bases that can never match: unknown letter, lower case, two letters, empty
    mat = build_em_matrix(ref, toy_phylo, ["1:N,2:a,3:AT,4:", "1:G, 2:T"], haps, make_args())
    want, _ = oracle_np.build_matrix_loops(ref, toy_phylo, ["1:N,2:a,3:AT,4:", "1:G, 2:T"], haps)
    assert np.array_equal(mat, want)
    with pytest.raises(ValueError):
        build_em_matrix(ref, toy_phylo, ["1:A", "2A"], haps, make_args())
    with pytest.raises(ValueError):
        build_em_matrix(ref, toy_phylo, ["1:A", ""], haps, make_args())
    with pytest.raises(KeyError):
        build_em_matrix(ref, toy_phylo, ["1:A", "77:A"], haps, make_args())
    with pytest.raises(KeyError):
        build_em_matrix(ref, toy_phylo, ["1:A"], haps + ["nope"], make_args())
    # the malformed row comes first -> ValueError wins over the later KeyError
    with pytest.raises(ValueError):
        build_em_matrix(ref, toy_phylo, ["1:A:C", "77:A"], haps, make_args())


def test_build_device_resident_and_argmax(phylo17):
    haps = sorted(phylo17.hap_var)
    mix = synth.make_mixture(phylo17, phylo17.refseq, [("H1", 0.7), ("L3e", 0.3)], 600, seed=2)
    dmat, cnt, ms = build_em_matrix_device(phylo17.refseq, phylo17, mix.signatures, haps,
                                           want_counts=True)
    host = build_em_matrix(phylo17.refseq, phylo17, mix.signatures, haps, make_args())
    assert dmat.shape == host.shape
    assert np.array_equal(dmat.to_host(), host)
    assert np.array_equal(dmat.argmax_rows(), np.argmax(host, 1))
    dmat.free()


def test_verbose_messages(toy_phylo, capsys):
    reads = ["1:A,2:T"] * 3
    build_em_matrix("AAAAAAAAA", toy_phylo, reads, list("ABCDEFGHI"), make_args(verbose=True))
    err = capsys.readouterr().err
    assert err.startswith("Building EM input matrix...\n") and err.endswith("Done.\n\n")


def test_build_class_kernel_equals_dense_kernel(phylo17, monkeypatch):
    """The class kernel (shared chains, sparse deviation lists) and the plain
    dense kernel (MXB_BUILD_DENSE=1) produce the same bits; noisy reads with
    unknown bases exercise non-reference observations."""
    haps = sorted(phylo17.hap_var)
    mix = synth.make_mixture(phylo17, phylo17.refseq,
                             [("H1", 0.4), ("L3e", 0.3), ("U5a1", 0.2), ("M7", 0.1)], 1500,
                             err=0.04, seed=21)
    tables = HapVarBaseMatrix(phylo17.refseq, phylo17, haps).pack()
    csr = mix.csr(tables)
    csr.base_code = csr.base_code.copy()
    csr.base_code[::37] = 255          # bases outside the alphabet never match
    mat, cnt, _, _ = build_matrix_from_csr(tables, csr, want_counts=True)
    mat_nc, _, _, _ = build_matrix_from_csr(tables, csr, want_counts=False)
    monkeypatch.setenv("MXB_BUILD_DENSE", "1")
    d_mat, d_cnt, _, _ = build_matrix_from_csr(tables, csr, want_counts=True)
    monkeypatch.delenv("MXB_BUILD_DENSE")
    o_mat, o_cnt = oracle_c.build_matrix(tables, csr)
    assert np.array_equal(mat, o_mat) and np.array_equal(cnt, o_cnt)
    assert np.array_equal(mat_nc, o_mat)
    assert np.array_equal(d_mat, o_mat) and np.array_equal(d_cnt, o_cnt)
    # every row through the large-pool launch only
    monkeypatch.setenv("MXB_BUILD_TIERS", "2")
    t_mat, t_cnt, _, _ = build_matrix_from_csr(tables, csr, want_counts=True)
    monkeypatch.delenv("MXB_BUILD_TIERS")
    assert np.array_equal(t_mat, o_mat) and np.array_equal(t_cnt, o_cnt)
    # deviation lists relative to the marker-free pattern instead of the majority pattern
    # (read when the tables are packed on the device)
    monkeypatch.setenv("MXB_BUILD_BASE_REF", "1")
    tables_ref = HapVarBaseMatrix(phylo17.refseq, phylo17, haps).pack()
    r_mat, r_cnt, _, _ = build_matrix_from_csr(tables_ref, csr, want_counts=True)
    monkeypatch.delenv("MXB_BUILD_BASE_REF")
    assert np.array_equal(r_mat, o_mat) and np.array_equal(r_cnt, o_cnt)


def test_build_unsorted_and_repeated_positions(phylo17):
    """Signature order is the summation order (preprocess.py:92-95): shuffled
    and repeated positions must give the reference's sums, not a sorted sum."""
    haps = sorted(phylo17.hap_var)
    positions = sorted(phylo17.variants)
    rs = np.random.RandomState(5)
    reads = []
    for _ in range(64):
        start = rs.randint(0, len(positions) - 120)
        pick = rs.permutation(np.arange(start, start + 120))[:rs.randint(1, 100)]
        pick = np.concatenate([pick, pick[:rs.randint(0, 4)]])     # repeats
        reads.append(",".join("%d:%s" % (positions[i], "ACGT"[rs.randint(4)]) for i in pick))
    mat = build_em_matrix(phylo17.refseq, phylo17, reads, haps, make_args())
    tables = HapVarBaseMatrix(phylo17.refseq, phylo17, haps).pack()
    csr, err = parse_signatures(reads, tables)
    assert err is None
    o_mat, _ = oracle_c.build_matrix(tables, csr)
    assert np.array_equal(mat, o_mat)
