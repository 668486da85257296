"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck /
initcheck / synccheck):  compute-sanitizer --tool racecheck python scripts/sanitize.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mixemt_b200  # noqa: E402
from mixemt_b200 import consumers, em, synth  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import DeviceMatrix, get_context  # noqa: E402
from conftest import make_args  # noqa: E402


def main():
    ctx = get_context()
    # kernel 1: class kernel (both tiers + dense rows) and dense kernel
    phylo, refseq = synth.synthetic_phylo(n_hap=1300, n_pos=500, ref_len=5000, seed=3)
    haps = sorted(phylo.hap_var)
    mix = synth.make_mixture(phylo, refseq, [(haps[5], 0.7), (haps[77], 0.3)], 600,
                             frag_len=900, seed=2)
    tables = HapVarBaseMatrix(refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    out, cnt, dmat, _ = build_matrix_from_csr(tables, csr, ctx=ctx, want_counts=True,
                                              keep_device=True)
    os.environ["MXB_BUILD_DENSE"] = "1"
    out2, cnt2, _, _ = build_matrix_from_csr(tables, csr, ctx=ctx, want_counts=True)
    del os.environ["MXB_BUILD_DENSE"]
    assert np.array_equal(out, out2) and np.array_equal(cnt, cnt2)
    # kernel 2: ring pass (single + pair), finish, general path, read_mix
    rs = np.random.RandomState(0)
    for n, h, n_multi in ((700, 1300, 1), (700, 1300, 3), (300, 5408, 2), (200, 40, 2), (64, 3, 1)):
        mat = -rs.gamma(2.0, 5.0, size=(n, h))
        wts = rs.randint(1, 9, size=n)
        inits = np.log(rs.dirichlet([1.0] * h, size=n_multi))
        dev = DeviceMatrix.from_host(ctx, mat)
        a = make_args(n_multi=n_multi, max_iter=12, tolerance=1e-9)
        props, read_mix, info, _ = em.run_em_device(dev, wts, a, inits=inits)
        assert np.isfinite(props).all() and np.isfinite(read_mix).all(), (n, h)
        dev.free()
    # kernel 2 over class tiles: blocks of identical columns (8 lanes per row), a block of noise
    # rows in the middle (all columns distinct: 16 warps per row), several work items per batch,
    # and a row that underflows in linear space (log-space rescue step)
    n, h = 420, 5408
    mat = np.repeat(-rs.gamma(2.0, 3.0, size=(n, h // 64 + 1)), 64, axis=1)[:, :h].copy()
    mat[130:170] -= rs.gamma(2.0, 1.0, size=(40, h))
    mat[:, 0] = 0.0
    mat[:8, 1:] -= 800.0
    init = np.full(h, 1.0)
    init[0] = 1e-300
    dev = DeviceMatrix.from_host(ctx, mat)
    a = make_args(n_multi=2, max_iter=6, tolerance=1e-9)
    inits = np.log(np.stack([init / init.sum(), rs.dirichlet([1.0] * h)]))
    props, read_mix, info, _ = em.run_em_device(dev, rs.randint(1, 9, size=n), a, inits=inits)
    assert np.isfinite(props).all() and np.isfinite(read_mix).all()
    dev.free()
    p, m = em.run_em(dmat, mix.weights, make_args(max_iter=10))
    # consumers
    votes, best = consumers.vote_count(m, mix.weights)
    assert votes.sum() == mix.weights.sum()
    consumers.assign_rows(m, p, [int(np.argmax(p)), 3, 9], 2.0)
    small = consumers.gather_columns(dmat, [1, 5, 7])
    assert small.shape == (mix.n_rows, 3)
    mat1 = np.zeros((3, 3))
    em.em_step(mat1, np.ones(3), np.log(np.ones(3) / 3), np.empty((3, 3)))
    ctx.trim()
    print("sanitize run ok")


if __name__ == "__main__":
    main()
