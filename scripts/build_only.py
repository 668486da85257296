"""Kernel 1 alone on the config-2 workload (ncu target / timing loop)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import load_workload  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import get_context  # noqa: E402


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    phylo, haps, mix = load_workload(frags, 2)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    ctx = get_context()
    times = []
    for _ in range(reps):
        _, _, dmat, ms = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=False,
                                               keep_device=True)
        times.append(ms)
        dmat.free()
    n, h = csr.n_rows, len(haps)
    best = min(times)
    print("build %d x %d: ms %s best %.3f -> %.1f Gcells/s, %.0f GB/s written"
          % (n, h, ["%.3f" % t for t in times], best, n * h / best / 1e6, n * h * 8 / best / 1e6))


if __name__ == "__main__":
    main()
