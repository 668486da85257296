"""Synthetic mtDNA mixtures of the shapes BASELINE.json names.

Workload generator for tests and bench.py, not part of the drop-in path.  Per
fragment: pick a source haplogroup by mixture weight, a uniform start in
[0, len(refseq) - L), observe every variant position inside the window with the
haplotype's expected base (its marker there, else the reference base) and
replace it by a uniformly random *other* base with probability ``err``.
Fragments are then reduced to unique signatures + multiplicities, which is
what the reference's process_reads/reduce_reads (preprocess.py:99-174) hand to
build_em_matrix; signature strings use the reference's format
(``"%d:%s"`` joined by ``,``, preprocess.py:142-148) and sort order
(plain string sort, preprocess.py:219).
"""
import numpy as np

from .preprocess import pos_from_var, der_allele, SignatureCSR

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


class Mixture(object):
    """Unique signatures of a synthetic mixture.

    ``pos_idx`` indexes ``sorted(phylo.variants)``; ``base_ascii`` holds the
    observed base letters; rows follow ``signatures`` (string-sorted, the order
    ``build_em_input`` hands them over in, preprocess.py:219) when those were
    requested, and the same grouping by first position otherwise."""

    def __init__(self, row_ptr, pos_idx, base_ascii, weights, signatures, n_fragments):
        self.row_ptr = row_ptr
        self.pos_idx = pos_idx
        self.base_ascii = base_ascii
        self.weights = weights
        self.signatures = signatures
        self.n_fragments = n_fragments

    @property
    def n_rows(self):
        return len(self.row_ptr) - 1

    def csr(self, tables):
        """CSR with symbol codes of a packed ``HapVarBaseMatrix``."""
        return SignatureCSR(self.row_ptr, self.pos_idx, tables.sym2code[self.base_ascii])


def expected_bases(phylo, refseq, hap, positions):
    """ASCII expected base of ``hap`` at every position (marker or reference,
    the rule of reference preprocess.py:56-67 / :75-83)."""
    exp = np.frombuffer("".join(refseq[p] for p in positions).encode("ascii"),
                        dtype=np.uint8).copy()
    index = {p: i for i, p in enumerate(positions)}
    for var in phylo.hap_var[hap]:
        pos = pos_from_var(var)
        der = der_allele(var)
        if pos in index and der != refseq[pos]:
            exp[index[pos]] = ord(der)
    return exp


def make_mixture(phylo, refseq, mixture, n_fragments, frag_len=300, err=0.002, seed=1,
                 strings=True):
    """``mixture``: list of ``(haplogroup id, fraction)``."""
    rs = np.random.RandomState(seed)
    positions = sorted(phylo.variants)
    pos_arr = np.asarray(positions, dtype=np.int64)
    haps = [h for h, _ in mixture]
    frac = np.asarray([f for _, f in mixture], dtype=np.float64)
    frac = frac / frac.sum()
    exp = [expected_bases(phylo, refseq, h, positions) for h in haps]

    src = rs.choice(len(haps), size=n_fragments, p=frac)
    start = rs.randint(0, len(refseq) - frag_len, size=n_fragments)
    lo = np.searchsorted(pos_arr, start, side="left")
    hi = np.searchsorted(pos_arr, start + frag_len, side="left")
    n_err = rs.binomial(hi - lo, err)

    counts = {}
    lo_l, hi_l, src_l, err_l = lo.tolist(), hi.tolist(), src.tolist(), n_err.tolist()
    for f in range(n_fragments):
        a, b = lo_l[f], hi_l[f]
        if b <= a:
            continue  # no variant site observed: the fragment never reaches reduce_reads
        obs = exp[src_l[f]][a:b]
        if err_l[f]:
            obs = obs.copy()
            for k in rs.choice(b - a, size=err_l[f], replace=False).tolist():
                others = _BASES[_BASES != obs[k]]
                obs[k] = others[rs.randint(len(others))]
        key = (a, obs.tobytes())
        counts[key] = counts.get(key, 0) + 1

    keys = list(counts)
    if strings:
        tok = {}

        def sig_of(key):
            a, raw = key
            parts = []
            for i, c in enumerate(raw):
                t = tok.get((a + i, c))
                if t is None:
                    t = tok[(a + i, c)] = "%d:%s" % (positions[a + i], chr(c))
                parts.append(t)
            return ",".join(parts)
        sigs = [sig_of(k) for k in keys]
        order = sorted(range(len(keys)), key=sigs.__getitem__)
        keys = [keys[i] for i in order]
        signatures = [sigs[i] for i in order]
    else:
        # no strings wanted: still the reference's grouping (rows that start at the same
        # variant position are neighbours, positions ordered as decimal strings)
        signatures = None
        rank = {p: r for r, p in enumerate(sorted(range(len(positions)),
                                                  key=lambda i: "%d:" % positions[i]))}
        keys.sort(key=lambda k: (rank[k[0]], k[1]))

    lens = np.fromiter((len(k[1]) for k in keys), dtype=np.int64, count=len(keys))
    row_ptr = np.zeros(len(keys) + 1, dtype=np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    base_ascii = np.frombuffer(b"".join(k[1] for k in keys), dtype=np.uint8).copy()
    starts = np.fromiter((k[0] for k in keys), dtype=np.int64, count=len(keys))
    pos_idx = (np.repeat(starts - row_ptr[:-1], lens) + np.arange(row_ptr[-1])).astype(np.int32)
    weights = np.fromiter((counts[k] for k in keys), dtype=np.int64, count=len(keys))
    return Mixture(row_ptr, pos_idx, base_ascii, weights, signatures, n_fragments)


def synthetic_phylo(n_hap=512, n_pos=400, ref_len=4000, markers_per_hap=12, seed=7):
    """A random haplotype table with Phylotree-like statistics (each haplotype
    carries a handful of derived bases; per-position mutation counts 1..50) for
    tests that must not depend on the Build 17 fixture.  Returns
    ``(PhyloTables, refseq)``."""
    import collections
    from .phylo_tables import PhyloTables
    rs = np.random.RandomState(seed)
    refseq = "".join("ACGT"[i] for i in rs.randint(0, 4, size=ref_len))
    positions = np.sort(rs.choice(ref_len, size=n_pos, replace=False))
    variants = {}
    for p in positions.tolist():
        cnt = collections.Counter()
        for _ in range(1 + int(rs.randint(0, 3) == 0)):
            others = [b for b in "ACGT" if b != refseq[p]]
            cnt[others[rs.randint(3)]] += int(min(60, 1 + rs.geometric(0.4)))
        variants[p] = cnt
    hap_var = {}
    seen = set()
    j = 0
    while len(hap_var) < n_hap:
        k = max(1, int(rs.poisson(markers_per_hap)))
        ps = np.sort(rs.choice(positions, size=min(k, n_pos), replace=False)).tolist()
        vs = []
        for p in ps:
            der = list(variants[p])[rs.randint(len(variants[p]))]
            vs.append("%s%d%s" % (refseq[p], p + 1, der))
        key = ",".join(vs)
        j += 1
        if key in seen:
            continue
        seen.add(key)
        hap_var["S%d" % j] = vs
    return PhyloTables(variants, hap_var, refseq), refseq


def random_rows(phylo, refseq, mixture, n_rows, frag_len=300, err=0.002, seed=1,
                reference_order=True):
    """``n_rows`` fragment signatures drawn like :func:`make_mixture` draws
    fragments, fully vectorised and *without* the reduction to unique signatures
    (duplicates stay, every weight is 1).  For shard-sized workloads (millions of
    rows per GPU, BASELINE.json config 3) where the Python dedupe loop of
    ``make_mixture`` would take minutes.  Rows are grouped like the reference's string-sorted
    signatures (``reference_order``) or left in generation order."""
    rs = np.random.RandomState(seed)
    positions = sorted(phylo.variants)
    pos_arr = np.asarray(positions, dtype=np.int64)
    haps = [h for h, _ in mixture]
    frac = np.asarray([f for _, f in mixture], dtype=np.float64)
    frac = frac / frac.sum()
    exp = np.stack([expected_bases(phylo, refseq, h, positions) for h in haps])
    chunks = []
    total_rows = 0
    while total_rows < n_rows:
        m = min(n_rows - total_rows + 1024, 2000000)
        src = rs.choice(len(haps), size=m, p=frac)
        start = rs.randint(0, len(refseq) - frag_len, size=m)
        lo = np.searchsorted(pos_arr, start, side="left")
        hi = np.searchsorted(pos_arr, start + frag_len, side="left")
        keep = hi > lo
        src, lo, hi = src[keep], lo[keep], hi[keep]
        take = min(len(src), n_rows - total_rows)
        src, lo, hi = src[:take], lo[:take], hi[:take]
        lens = (hi - lo).astype(np.int64)
        ptr = np.zeros(take + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        idx = (np.repeat(lo - ptr[:-1], lens) + np.arange(ptr[-1])).astype(np.int32)
        base = exp[np.repeat(src, lens), idx]
        flip = np.nonzero(rs.rand(len(base)) < err)[0]
        if len(flip):
            code = np.searchsorted(_BASES, base[flip])
            base[flip] = _BASES[(code + rs.randint(1, 4, size=len(flip))) % 4]
        chunks.append((lens, idx, base))
        total_rows += take
    lens = np.concatenate([c[0] for c in chunks])
    row_ptr = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    pos_idx = np.concatenate([c[1] for c in chunks])
    base = np.concatenate([c[2] for c in chunks])
    if reference_order:
        # build_em_input hands the rows over string-sorted (preprocess.py:219): rows that
        # start at the same variant position are neighbours, and positions follow each other
        # as the strings "pos:" ("1000:" < "1001:" < "100:" < "101:").  Same grouping here, without
        # building ten million strings: order by the string rank of the first position, then
        # by the window's end.
        rank = np.empty(len(positions), dtype=np.int64)
        rank[np.argsort(np.array(["%d:" % p for p in positions]))] = np.arange(len(positions))
        first = pos_idx[row_ptr[:-1]]
        order = np.lexsort((lens, rank[first]))
        new_ptr = np.zeros_like(row_ptr)
        np.cumsum(lens[order], out=new_ptr[1:])
        take = (np.repeat(row_ptr[:-1][order] - new_ptr[:-1], lens[order])
                + np.arange(new_ptr[-1]))
        row_ptr, pos_idx, base = new_ptr, pos_idx[take], base[take]
    return Mixture(row_ptr, pos_idx, base, np.ones(len(lens), dtype=np.int64), None, len(lens))
