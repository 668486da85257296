"""Class-tile statistics of the config-2 matrix (analysis only; lives under tests/ because it
builds the matrix with the CPU oracle).

Per batch size: classes of bit-identical columns per batch, tile bytes, team widths, the
work per width class, and the class counts of super-batches (common refinement of G
consecutive batches).  Saves the per-batch class counts to /tmp/tile_plan.npz."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from mixemt_b200.preprocess import HapVarBaseMatrix
from oracle import oracle_c


def class_ids(block, rnd):
    """Class id per column of a row block (columns equal in every row): a random fp64
    projection separates distinct columns with overwhelming probability."""
    key = rnd[:block.shape[0]] @ block
    _, inv = np.unique(key, return_inverse=True)
    return inv


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    t0 = time.time()
    phylo, haps, mix = bench.load_workload(frags, 2)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    n = csr.n_rows
    print("rows", n, "gen %.1fs" % (time.time() - t0), flush=True)
    mat, _ = oracle_c.build_matrix(tables, csr, want_counts=False)
    print("matrix", mat.shape, "%.1fs" % (time.time() - t0), flush=True)
    h = mat.shape[1]
    rs = np.random.RandomState(5)
    rnd = rs.rand(4096) + 0.5
    out = {}
    for rows in (32, 64, 128, 256, 1024):
        nb = (n + rows - 1) // rows
        ncls = np.empty(nb, dtype=np.int64)
        for b in range(nb):
            ncls[b] = class_ids(mat[b * rows:(b + 1) * rows], rnd).max() + 1
        nr = np.minimum(rows, n - np.arange(nb) * rows)
        pad = ((ncls + 63) // 64) * 64
        print("batch %4d rows: %5d batches, classes median %d mean %.0f p90 %d p99 %d max %d; "
              "tiles %.3f GB (padded to 64: %.3f GB), maps %.1f MB"
              % (rows, nb, np.median(ncls), ncls.mean(), np.percentile(ncls, 90),
                 np.percentile(ncls, 99), ncls.max(), (ncls * nr).sum() * 8 / 1e9,
                 (pad * nr).sum() * 8 / 1e9, nb * h * 2 / 1e6), flush=True)
        out["ncls_%d" % rows] = ncls
        if rows in (32, 64, 128):
            tw = np.ones(nb, dtype=np.int64)
            for k in range(4):
                tw = np.where(ncls > 512 * tw, tw * 2, tw)
            nk = np.maximum(1, -(-ncls // (64 * tw)))
            cpad = 64 * tw * nk
            for t in (1, 2, 4, 8, 16):
                sel = tw == t
                if sel.any():
                    print("    tw %2d: %5d batches, %.3f GB of tiles (%.1f %%)"
                          % (t, sel.sum(), (cpad[sel] * nr[sel]).sum() * 8 / 1e9,
                             100.0 * (cpad[sel] * nr[sel]).sum() / (cpad * nr).sum()), flush=True)
    np.savez("/tmp/tile_plan.npz", **out)


if __name__ == "__main__":
    main()
