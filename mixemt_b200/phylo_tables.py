"""A light stand-in for ``mixemt.phylotree.Phylotree`` holding only what the
hot path reads: ``variants`` (position -> {derived base: count}) and ``hap_var``
(haplogroup id -> list of SNP strings), see reference phylotree.py:25-31.

The reference's Phylotree class (CSV parsing, tree filters) is out of scope and
is used unchanged when mixemt itself is installed.  This container exists so
that tests and the benchmark can run on a box where ``/root/reference`` is not
present, from tables that ``oracle/make_golden.py`` exported with the
reference's own Phylotree (tests/golden/phylotree17*.npz).
"""
import collections

import numpy as np


class PhyloTables(object):
    def __init__(self, variants, hap_var, refseq=None, meta=None):
        self.variants = variants
        self.hap_var = hap_var
        self.refseq = refseq
        self.meta = meta or {}

    def get_variant_pos(self):
        return sorted(self.variants)

    @classmethod
    def from_phylotree(cls, phylo, refseq=None, meta=None):
        variants = {pos: collections.Counter(cnt) for pos, cnt in phylo.variants.items()}
        hap_var = {hap: list(vs) for hap, vs in phylo.hap_var.items()}
        return cls(variants, hap_var, refseq, meta)

    # -- npz round trip ----------------------------------------------------------
    def save(self, path):
        haps = list(self.hap_var)
        pos, base, count = [], [], []
        for p in sorted(self.variants):
            for b, c in sorted(self.variants[p].items()):
                pos.append(p)
                base.append(b)
                count.append(c)
        np.savez_compressed(
            path,
            hap_ids=np.array("\n".join(haps)),
            hap_vars=np.array("\n".join(",".join(self.hap_var[h]) for h in haps)),
            var_pos=np.asarray(pos, dtype=np.int32),
            var_base=np.array("".join(base)),
            var_count=np.asarray(count, dtype=np.int32),
            refseq=np.array(self.refseq or ""),
            meta=np.array(repr(sorted(self.meta.items()))))

    @classmethod
    def load(cls, path):
        with np.load(path, allow_pickle=False) as z:
            haps = str(z["hap_ids"]).split("\n")
            var_lines = str(z["hap_vars"]).split("\n")
            if len(var_lines) != len(haps):
                raise ValueError("corrupt phylo table fixture %s" % path)
            hap_var = {h: ([v for v in line.split(",")] if line else [])
                       for h, line in zip(haps, var_lines)}
            variants = collections.defaultdict(collections.Counter)
            bases = str(z["var_base"])
            for p, b, c in zip(z["var_pos"].tolist(), bases, z["var_count"].tolist()):
                variants[p][b] += c
            refseq = str(z["refseq"]) or None
            meta = str(z["meta"])
        return cls(dict(variants), hap_var, refseq, {"repr": meta})
