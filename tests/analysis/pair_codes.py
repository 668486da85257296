"""Distinct (cell 2c, cell 2c+1) value pairs per row of the config-2 matrix, for 2- and 4-cell
chunks (analysis only; lives under tests/ because it builds its sample with the CPU oracle): would a dictionary of chunk values fit 256 entries?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from row_runs import sample_matrix

if __name__ == "__main__":
    phylo, haps, mat = sample_matrix(int(sys.argv[1]) if len(sys.argv) > 1 else 1500)
    bits = np.ascontiguousarray(mat).view(np.uint64)
    n, h = bits.shape
    d1 = np.array([len(np.unique(r)) for r in bits])
    for w in (2, 4):
        hh = (h // w) * w
        cnt = np.empty(n, dtype=np.int64)
        for i in range(n):
            ch = bits[i, :hh].reshape(-1, w)
            cnt[i] = len(np.unique(ch, axis=0))
        print("chunk of %d cells: distinct chunks per row median %d p90 %d p99 %d; <=256: %.3f "
              "(single values <=256: %.3f)" % (w, np.median(cnt), np.percentile(cnt, 90),
                                               np.percentile(cnt, 99), (cnt <= 256).mean(),
                                               (d1 <= 256).mean()))
