"""Device-side consumers of the EM result (mixemt_b200.consumers, SURVEY.md 8f
N2-N4) against the reference's outputs (tests/golden/golden_consumers.npz) and
the CPU oracle, through the C-ABI."""
import os

import numpy as np
import pytest

import mixemt_b200
from mixemt_b200 import consumers
from mixemt_b200.runtime import DeviceMatrix, get_context
from oracle import oracle_np
from conftest import load_golden, make_args

pytestmark = pytest.mark.gpu


def _decode_assign(res, n):
    out = np.full(n, -2, dtype=np.int32)
    for name, rows in res.items():
        out[sorted(rows)] = -1 if name == 'unassigned' else int(name[3:]) - 1
    return out


@pytest.mark.parametrize("seed", [71, 72])
@pytest.mark.parametrize("on_device", [False, True])
def test_consumers_match_reference(seed, on_device):
    g = load_golden("golden_consumers.npz")
    props, mix, wts = oracle_np.synthetic_em_result(seed)
    haps = ["hg%d" % j for j in range(mix.shape[1])]
    arg = DeviceMatrix.from_host(get_context(), mix) if on_device else mix
    for min_reads in (1, 60, 400):
        got = consumers.find_contribs_from_reads(arg, wts, make_args(min_reads=min_reads))
        assert got == g["s%d_contribs_r%d" % (seed, min_reads)].tolist()
    assert np.array_equal(consumers.read_votes(arg), np.argmax(mix, 1))
    top = g["s%d_top" % seed].tolist()
    contribs = [["hap%d" % (k + 1), haps[j], props[j]] for k, j in enumerate(top)]
    for tag, fold, cons in (("f2", 2.0, contribs), ("f1p2", 1.2, contribs[:2]),
                            ("single", 2.0, contribs[:1])):
        res = consumers.assign_read_indexes(cons, (props, arg), haps, list(range(len(mix))), fold)
        assert np.array_equal(_decode_assign(res, len(mix)), g["s%d_assign_%s" % (seed, tag)])
        assert all(isinstance(v, set) for v in res.values())
    red, new_haps = consumers.reduce_em_matrix(arg, haps, contribs)
    assert [haps.index(x) for x in new_haps] == g["s%d_reduced_cols" % seed].tolist()
    red_host = red.to_host() if on_device else red
    assert isinstance(red, DeviceMatrix) == on_device
    assert np.array_equal(red_host, g["s%d_reduced" % seed])


def test_votes_at_build17_width_and_edge_rows():
    rs = np.random.RandomState(5)
    n, h = 3001, 5408
    mix = -rs.gamma(2.0, 5.0, size=(n, h))
    mix[rs.rand(n, h) < 0.2] = 0.0
    mix[0] = -np.inf                      # all -inf: argmax 0
    mix[1] = 0.0                          # all equal: argmax 0
    mix[2, -1] = 1.0                      # maximum in the last column
    wts = rs.randint(1, 1000, size=n).astype(np.int64)
    votes, best = consumers.vote_count(mix, wts)
    assert np.array_equal(best, np.argmax(mix, 1))
    assert np.array_equal(votes, np.bincount(best, weights=wts, minlength=h).astype(np.int64))
    got = consumers.find_contribs_from_reads(mix, wts, make_args(min_reads=2000))
    assert got == oracle_np.find_contribs_from_reads(mix, wts, 2000)


def test_refinement_stays_on_device(phylo17):
    """bin/mixemt:311-320: reduce_em_matrix then run_em on the column subset,
    here without the matrices leaving HBM; must equal the host-array route."""
    from mixemt_b200 import synth, em
    from mixemt_b200.preprocess import build_em_matrix_device
    haps = sorted(phylo17.hap_var)
    mix = synth.make_mixture(phylo17, phylo17.refseq, [("H1", 0.6), ("L3e", 0.4)], 3000, seed=8)
    dmat, _, _ = build_em_matrix_device(phylo17.refseq, phylo17, mix.signatures, haps)
    contribs = [["hap1", "H1", 0.6], ["hap2", "L3e", 0.4], ["hap3", "U5a1", 0.0]]
    small, new_haps = consumers.reduce_em_matrix(dmat, haps, contribs)
    assert new_haps == sorted(["H1", "L3e", "U5a1"]) and small.shape == (mix.n_rows, 3)
    host = dmat.to_host()
    want = host[:, [haps.index(x) for x in new_haps]]
    assert np.array_equal(small.to_host(), want)
    args = make_args(tolerance=1e-8)
    np.random.seed(3)
    p_dev, m_dev = mixemt_b200.run_em(small, mix.weights, args)
    np.random.seed(3)
    inits = np.log(np.random.dirichlet([1.0] * 3))[None, :]
    p_ref, m_ref, _ = oracle_np.run_em(want, mix.weights, inits, args.max_iter, args.tolerance)
    assert np.abs(p_dev - p_ref).max() < 1e-9
    assert np.abs(m_dev - m_ref).max() < 1e-9
    assert abs(p_dev[new_haps.index("H1")] - 0.6) < 0.05


def test_resident_results_feed_consumers(phylo17):
    """Resident mode: run_em's read-only result keeps its HBM copy and the
    consumers use it (no upload); a writable array is uploaded instead."""
    from mixemt_b200 import synth, runtime
    haps = sorted(phylo17.hap_var)[:700] + ["H1", "L3e"]
    haps = sorted(set(haps))
    mix = synth.make_mixture(phylo17, phylo17.refseq, [("H1", 0.5), ("L3e", 0.5)], 1500, seed=9)
    args = make_args(b200_resident=True, max_iter=50, min_reads=10)
    mat = mixemt_b200.build_em_matrix(phylo17.refseq, phylo17, mix.signatures, haps, args)
    np.random.seed(4)
    props, read_mix = mixemt_b200.run_em(mat, mix.weights, args)
    assert not read_mix.flags.writeable and runtime.lookup_resident(read_mix) is not None
    ctx = get_context()
    before = ctx.launch_count
    cons = consumers.find_contribs_from_reads(read_mix, mix.weights, args)
    assert cons == oracle_np.find_contribs_from_reads(read_mix, mix.weights, 10)
    assert ctx.launch_count - before == 2      # argmax + vote kernels, no upload path
    red, new_haps = consumers.reduce_em_matrix(mat, haps, [["hap1", "H1", 0.5], ["hap2", "L3e", 0.5]])
    assert np.array_equal(red, mat[:, [haps.index("H1"), haps.index("L3e")]])
    assert runtime.lookup_resident(red) is not None


def test_npy_streaming_roundtrip(tmp_path):
    rs = np.random.RandomState(2)
    mat = rs.randn(1237, 611)
    mat[5, 7] = -np.inf
    dev = DeviceMatrix.from_host(get_context(), mat)
    old = consumers._CHUNK_BYTES
    consumers._CHUNK_BYTES = 1 << 20           # several chunks, ragged last one
    try:
        path = consumers.save_npy(dev, str(tmp_path / "x.em"))
        assert path.endswith("x.em.npy")
        ref_path = str(tmp_path / "ref.npy")
        np.save(ref_path, mat)
        assert open(path, "rb").read() == open(ref_path, "rb").read()   # byte-identical file
        back = consumers.load_npy(ref_path)
        assert np.array_equal(back.to_host(), mat)
    finally:
        consumers._CHUNK_BYTES = old
    with pytest.raises(ValueError):
        np.save(str(tmp_path / "bad.npy"), mat.astype(np.float32))
        consumers.load_npy(str(tmp_path / "bad.npy"))


def test_consumer_argument_errors():
    ctx = get_context()
    dev = DeviceMatrix.from_host(ctx, np.zeros((4, 3)))
    with pytest.raises(mixemt_b200._lib.MixemtB200Error):
        consumers.gather_columns(dev, [0, 3])
    with pytest.raises(ValueError):
        consumers.vote_count(dev, [1, 2])
    with pytest.raises(IndexError):                      # like props[5] in the reference
        consumers.assign_rows(dev, np.ones(3) / 3, [0, 5], 2.0)
    with pytest.raises(mixemt_b200._lib.MixemtB200Error):
        consumers.assign_rows(dev, np.ones(6) / 6, [0, 5], 2.0)
