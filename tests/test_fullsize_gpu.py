"""Parity at BASELINE.json's full sizes (configs 2 and 5: ~140k-200k signature rows x
all Phylotree-17 haplotypes).  The CPU oracle (C + OpenMP) is fast enough to check the
matrix bit for bit on row samples spread over the whole matrix and to follow a few EM
iterations over ALL cells; everything else uses size-independent properties."""
import numpy as np
import pytest

from mixemt_b200 import em, synth
from mixemt_b200.preprocess import HapVarBaseMatrix, SignatureCSR, build_matrix_from_csr
from mixemt_b200.runtime import get_context
from oracle import oracle_c
from conftest import make_args

pytestmark = pytest.mark.gpu

N_ROWS = 140000


def _rows(csr, rows):
    lens = np.diff(csr.row_ptr)[rows]
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    idx = np.concatenate([np.arange(csr.row_ptr[r], csr.row_ptr[r + 1]) for r in rows])
    return SignatureCSR(ptr, csr.pos_idx[idx], csr.base_code[idx])


@pytest.mark.parametrize("fixture,mixture", [
    ("phylo17", [("H1", 0.5), ("L3e", 0.3), ("U5a1", 0.2)]),
    ("phylo17_cfg5", None),     # -U, --exclude_pos (doubled counts), custom haplotypes
])
def test_build_full_size(fixture, mixture, request, monkeypatch):
    phylo = request.getfixturevalue(fixture)
    haps = sorted(phylo.hap_var)
    if mixture is None:
        h1 = [h for h in haps if "H1" in h.split("/")][0]
        u5 = [h for h in haps if "U5a1" in h.split("/")][0]
        mixture = [(h1, 0.5), ("custom_hap2", 0.3), (u5, 0.2)]
    mix = synth.random_rows(phylo, phylo.refseq, mixture, N_ROWS, err=0.004, seed=17)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    ctx = get_context()
    _, _, dev, _ = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=False, keep_device=True)
    # class kernel == dense kernel over every cell (compared on the device side via host copies
    # of row blocks, 8192 rows at a time)
    monkeypatch.setenv("MXB_BUILD_DENSE", "1")
    _, _, dense, _ = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=False, keep_device=True)
    monkeypatch.delenv("MXB_BUILD_DENSE")
    from mixemt_b200._lib import lib, check, ptr
    a = np.empty((8192, len(haps)))
    b = np.empty((8192, len(haps)))
    for r0 in range(0, N_ROWS, 8192):
        k = min(8192, N_ROWS - r0)
        check(lib.mxb_matrix_download_rows(ctx.handle, dev.handle, r0, k, ptr(a)))
        check(lib.mxb_matrix_download_rows(ctx.handle, dense.handle, r0, k, ptr(b)))
        assert np.array_equal(a[:k], b[:k]), "class and dense kernels differ in rows %d.." % r0
    dense.free()
    # oracle on 400 rows spread over the matrix: bit-exact cells and match counts
    rows = np.unique(np.linspace(0, N_ROWS - 1, 400).astype(np.int64))
    sub = _rows(csr, rows)
    o_mat, o_cnt = oracle_c.build_matrix(tables, sub)
    g_mat, g_cnt, _, _ = build_matrix_from_csr(tables, sub, ctx=ctx, want_counts=True)
    assert np.array_equal(g_mat, o_mat) and np.array_equal(g_cnt, o_cnt)
    for i, r in enumerate(rows[::40]):
        check(lib.mxb_matrix_download_rows(ctx.handle, dev.handle, int(r), 1, ptr(a)))
        assert np.array_equal(a[0], o_mat[i * 40])
    dev.free()


def test_em_full_size_follows_the_oracle(phylo17):
    """Five EM iterations over all 7.6e8 cells of a config-2-sized matrix: proportions,
    row argmax and a sample of log responsibilities against the CPU oracle."""
    haps = sorted(phylo17.hap_var)
    mix = synth.random_rows(phylo17, phylo17.refseq, [("H1", 0.5), ("L3e", 0.3), ("U5a1", 0.2)],
                            N_ROWS, seed=23)
    rs = np.random.RandomState(4)
    wts = rs.randint(1, 60, size=N_ROWS).astype(np.float64)
    tables = HapVarBaseMatrix(phylo17.refseq, phylo17, haps).pack()
    ctx = get_context()
    host, _, dev, _ = build_matrix_from_csr(tables, mix.csr(tables), ctx=ctx, keep_device=True)
    inits = np.log(rs.dirichlet([1.0] * len(haps), size=1))
    a = make_args(max_iter=5, tolerance=1e-12)
    props, _, info, mix_dev = em.run_em_device(dev, wts, a, inits=inits, keep_device=True,
                                               want_host=False)
    o_props, o_mix, o_iters = oracle_c.run_em(host, wts, inits, 5, 1e-12)
    assert info["iterations"] == list(o_iters) == [5]
    assert np.abs(props - o_props).max() < 1e-12
    assert abs(props.sum() - 1.0) < 1e-12
    assert np.array_equal(mix_dev.argmax_rows(), np.argmax(o_mix, 1))
    from mixemt_b200._lib import lib, check, ptr
    row = np.empty((1, len(haps)))
    for r in (0, 1, N_ROWS // 2, N_ROWS - 1):
        check(lib.mxb_matrix_download_rows(ctx.handle, mix_dev.handle, r, 1, ptr(row)))
        assert np.abs(row[0] - o_mix[r]).max() < 1e-9
    dev.free()
    mix_dev.free()
