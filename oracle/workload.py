"""Synthetic workload for the CPU reference arm of bench.py.

TEST INFRASTRUCTURE ONLY.  ``bench.py --impl reference`` must not load the
product library, so it cannot use ``mixemt_b200.synth``; this module draws
fragments with the same recipe (source by mixture weight, uniform
start, every variant position of the window observed with the haplotype's
expected base, a random other base with probability ``err``) as plain
``{read_id: {pos: base}}`` dictionaries -- the input of the reference's own
``reduce_reads`` (preprocess.py:163-174) -- using only the reference's Phylotree
object and numpy.
"""
import numpy as np

from . import oracle_np

CONFIG2_MIXTURE = [("H1", 0.5), ("L3e", 0.3), ("U5a1", 0.2)]


def load_build17():
    """(phylo, refseq) of Phylotree Build 17 + RSRS through the reference's own
    Phylotree class (default options, bin/mixemt:86-88)."""
    from . import refload
    phylotree, _, _ = refload.load()
    csv, fa = refload.build17_paths()
    refseq = refload.read_fasta(fa)
    with open(csv) as handle:
        return phylotree.Phylotree(handle, refseq=refseq, anon_haps=True), refseq


def draw_read_obs(phylo, refseq, mixture, n_fragments, frag_len=300, err=0.002, seed=1):
    rs = np.random.RandomState(seed)
    positions = np.asarray(sorted(phylo.variants), dtype=np.int64)
    markers = oracle_np.marker_table(phylo, refseq)
    frac = np.asarray([f for _, f in mixture], dtype=np.float64)
    frac /= frac.sum()
    expected = []
    for hap, _ in mixture:
        table = markers[hap]
        expected.append([table.get(p, refseq[p]) for p in positions.tolist()])
    src = rs.choice(len(mixture), size=n_fragments, p=frac).tolist()
    start = rs.randint(0, len(refseq) - frag_len, size=n_fragments)
    lo = np.searchsorted(positions, start, side="left").tolist()
    hi = np.searchsorted(positions, start + frag_len, side="left").tolist()
    pos_list = positions.tolist()
    read_obs = {}
    for f in range(n_fragments):
        if hi[f] <= lo[f]:
            continue
        exp = expected[src[f]]
        obs = {}
        flips = rs.rand(hi[f] - lo[f]) < err
        for k in range(lo[f], hi[f]):
            base = exp[k]
            if flips[k - lo[f]]:
                others = [b for b in "ACGT" if b != base]
                base = others[rs.randint(len(others))]
            obs[pos_list[k]] = base
        read_obs["frag%07d" % f] = obs
    return read_obs


def sample_matrix(n_rows, seed=2, mixture=None):
    """``n_rows`` signature rows of the config-2 mixture x all Build-17
    haplogroups: fragments drawn as above, reduced by the REFERENCE's
    ``reduce_reads``, ordered like ``build_em_input`` (preprocess.py:218-220),
    matrix built by the vectorised restatement ``oracle_np.build_matrix_fast``
    (bit-identical to ``build_em_matrix``, tests/test_oracle.py; the reference's
    own loop needs ~95 ms per row).  Returns
    ``(phylo, refseq, reads, haplogroups, weights, matrix)``."""
    from . import refload
    _, preprocess, _ = refload.load()
    phylo, refseq = load_build17()
    read_obs = draw_read_obs(phylo, refseq, mixture or CONFIG2_MIXTURE,
                             max(2000, int(n_rows * 1.25) + 64), seed=seed)
    read_sigs = preprocess.reduce_reads(read_obs)
    reads = sorted(read_sigs)
    keep = np.sort(np.random.RandomState(seed).choice(len(reads), size=min(n_rows, len(reads)),
                                                      replace=False))
    reads = [reads[i] for i in keep.tolist()]
    weights = np.array([len(read_sigs[r]) for r in reads])
    haplogroups = sorted(phylo.hap_var)
    mat, _ = oracle_np.build_matrix_fast(refseq, phylo, reads, haplogroups)
    return phylo, refseq, reads, haplogroups, weights, mat
