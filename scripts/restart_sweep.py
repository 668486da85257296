"""Restart sweep (BASELINE.json config 4) on the resident config-2 matrix: wall clock of
run_em with n_multi restarts of a fixed number of iterations, two restarts per pass
(default) against one at a time (MXB_EM_NO_BATCH=1)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import load_workload  # noqa: E402
from mixemt_b200 import em  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import get_context  # noqa: E402


def main():
    n_multi = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    frags = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
    phylo, haps, mix = load_workload(frags, 2)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    ctx = get_context()
    _, _, dmat, _ = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=False, keep_device=True)
    n, h = dmat.shape
    inits = np.log(np.random.RandomState(1).dirichlet([1.0] * h, size=n_multi))
    args = argparse.Namespace(verbose=False, init_alpha=1.0, tolerance=-1.0, max_iter=iters,
                              n_multi=n_multi)
    res = {}
    for mode in ("batched", "sequential", "batched"):
        if mode == "sequential":
            os.environ["MXB_EM_NO_BATCH"] = "1"
        else:
            os.environ.pop("MXB_EM_NO_BATCH", None)
        ctx.synchronize()
        t0 = time.perf_counter()
        props, _, info, _ = em.run_em_device(dmat, mix.weights, args, want_host=False, inits=inits)
        dt = time.perf_counter() - t0
        total_iters = sum(info["iterations"])
        res[mode] = (dt, props)
        print("%-10s n_multi=%d: %.3f s, %d restart-iterations -> %.1f restart-iterations/s, "
              "%.3g cell-updates/s" % (mode, n_multi, dt, total_iters, total_iters / dt,
                                       total_iters * n * h / dt), flush=True)
    print("max |props batched - sequential| = %.3g" %
          np.abs(res["batched"][1] - res["sequential"][1]).max())


if __name__ == "__main__":
    main()
