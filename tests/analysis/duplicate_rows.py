"""How many rows of L = exp(M - rowmax) of the config-2 matrix are identical to another row
(analysis only; lives under tests/ because it builds the matrix with the CPU oracle)?  Identical
rows could be merged for the EM iterations (their weights add), exactly."""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from mixemt_b200.preprocess import HapVarBaseMatrix, SignatureCSR
from oracle import oracle_c


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    phylo, haps, mix = bench.load_workload(frags, 2)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    n = csr.n_rows
    seen = {}
    dup = 0
    dup_m = 0
    seen_m = set()
    step = 8192
    t0 = time.time()
    for a in range(0, n, step):
        b = min(n, a + step)
        rp = (csr.row_ptr[a:b + 1] - csr.row_ptr[a]).astype(np.int64)
        sl = slice(int(csr.row_ptr[a]), int(csr.row_ptr[b]))
        sub = SignatureCSR(rp, csr.pos_idx[sl].copy(), np.ascontiguousarray(csr.base_code[sl]))
        mat, _ = oracle_c.build_matrix(tables, sub, want_counts=False)
        lin = np.exp(mat - mat.max(axis=1, keepdims=True))
        for i in range(b - a):
            h = hashlib.blake2b(lin[i].tobytes(), digest_size=16).digest()
            if h in seen:
                dup += 1
            else:
                seen[h] = a + i
            hm = hashlib.blake2b(mat[i].tobytes(), digest_size=16).digest()
            if hm in seen_m:
                dup_m += 1
            else:
                seen_m.add(hm)
        print("rows %d / %d: %d duplicates of L so far (%d of M), %.0f s"
              % (b, n, dup, dup_m, time.time() - t0), flush=True)
    print("rows %d, distinct L rows %d (%.1f %% duplicates), distinct M rows %d"
          % (n, n - dup, 100.0 * dup / n, n - dup_m))


if __name__ == "__main__":
    main()
