"""Row structure of the config-2 matrix on a row sample (analysis only; lives under tests/ because it builds its sample with the CPU oracle).

How many cells of a row differ from the most common value of their column group, for several
group widths -- the size of a (group value + exception list) representation of L."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from mixemt_b200.preprocess import HapVarBaseMatrix, SignatureCSR
from oracle import oracle_c

def main():
    nsamp = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    frags = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    t0 = time.time()
    phylo, haps, mix = bench.load_workload(frags, 2)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    n = csr.n_rows
    print("rows", n, "gen %.1fs" % (time.time() - t0))
    rs = np.random.RandomState(0)
    pick = np.sort(rs.choice(n, size=min(nsamp, n), replace=False))
    lens = np.diff(csr.row_ptr)[pick]
    rp = np.zeros(len(pick) + 1, dtype=np.int64); np.cumsum(lens, out=rp[1:])
    idx = np.concatenate([np.arange(csr.row_ptr[r], csr.row_ptr[r + 1]) for r in pick])
    sub = SignatureCSR(rp, csr.pos_idx[idx].copy(), np.ascontiguousarray(csr.base_code[idx]))
    mat, _ = oracle_c.build_matrix(tables, sub, want_counts=False)
    h = mat.shape[1]
    print("sample", mat.shape)
    bits = mat.view(np.uint64)
    distinct = np.array([len(np.unique(r)) for r in bits])
    print("distinct values per row: median %d mean %.0f p90 %d p99 %d max %d; <=256: %.3f"
          % (np.median(distinct), distinct.mean(), np.percentile(distinct, 90),
             np.percentile(distinct, 99), distinct.max(), (distinct <= 256).mean()))
    # row-global majority
    e0 = np.empty(len(bits), dtype=np.int64)
    for i, r in enumerate(bits):
        _, c = np.unique(r, return_counts=True)
        e0[i] = h - c.max()
    print("cells != row majority: median %d mean %.0f p90 %d p99 %d max %d"
          % (np.median(e0), e0.mean(), np.percentile(e0, 90), np.percentile(e0, 99), e0.max()))
    for gw in (8, 16, 32, 64, 128):
        ng = (h + gw - 1) // gw
        pad = ng * gw - h
        exc = np.zeros(len(bits), dtype=np.int64)
        nonuni = np.zeros(len(bits), dtype=np.int64)
        for i, r in enumerate(bits):
            rr = np.concatenate([r, np.full(pad, r[-1], dtype=np.uint64)]).reshape(ng, gw)
            s = np.sort(rr, axis=1)
            # longest run of equal values per group
            best = np.ones(ng, dtype=np.int64); run = np.ones(ng, dtype=np.int64)
            for k in range(1, gw):
                same = s[:, k] == s[:, k - 1]
                run = np.where(same, run + 1, 1)
                best = np.maximum(best, run)
            exc[i] = (gw - best).sum()
            nonuni[i] = (best < gw).sum()
        byt = ng * 8 + exc * 10
        print("group %3d: groups %4d  exceptions/row median %d mean %.0f p90 %d p99 %d max %d | "
              "non-uniform groups mean %.0f | bytes/row (8B group values + 10B/exception) mean %.0f"
              % (gw, ng, np.median(exc), exc.mean(), np.percentile(exc, 90), np.percentile(exc, 99),
                 exc.max(), nonuni.mean(), byt.mean()))

if __name__ == "__main__":
    main()
