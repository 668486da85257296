"""Run-length structure of the rows of the config-2 matrix under different column orders
(analysis only; lives under tests/ because it builds its sample with the CPU oracle)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from mixemt_b200.preprocess import HapVarBaseMatrix, SignatureCSR
from oracle import oracle_c

def sample_matrix(nsamp, frags=1000000, seed=2):
    phylo, haps, mix = bench.load_workload(frags, seed)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    n = csr.n_rows
    rs = np.random.RandomState(0)
    pick = np.sort(rs.choice(n, size=min(nsamp, n), replace=False))
    lens = np.diff(csr.row_ptr)[pick]
    rp = np.zeros(len(pick) + 1, dtype=np.int64); np.cumsum(lens, out=rp[1:])
    idx = np.concatenate([np.arange(csr.row_ptr[r], csr.row_ptr[r + 1]) for r in pick])
    sub = SignatureCSR(rp, csr.pos_idx[idx].copy(), np.ascontiguousarray(csr.base_code[idx]))
    mat, _ = oracle_c.build_matrix(tables, sub, want_counts=False)
    return phylo, haps, mat

def runs(bits):
    return 1 + (bits[:, 1:] != bits[:, :-1]).sum(axis=1)

def report(name, r):
    print("%-28s runs/row median %d mean %.0f p90 %d p99 %d max %d"
          % (name, np.median(r), r.mean(), np.percentile(r, 90), np.percentile(r, 99), r.max()))

if __name__ == "__main__":
    phylo, haps, mat = sample_matrix(int(sys.argv[1]) if len(sys.argv) > 1 else 1500)
    bits = np.ascontiguousarray(mat).view(np.uint64)
    report("name-sorted (as built)", runs(bits))
    tree_order = list(phylo.hap_var)
    col = {h: i for i, h in enumerate(haps)}
    perm = np.array([col[h] for h in tree_order])
    report("hap_var insertion order", runs(bits[:, perm]))
