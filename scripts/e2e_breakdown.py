"""Stage-by-stage wall clock of the drop-in run_em(host ndarray) call on config 2
(diagnostic; MXB_TIMING=1 makes the native library print its own stages)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MXB_TIMING", "1")

import torch  # noqa: E402
import mixemt_b200  # noqa: E402
from bench import load_workload  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import DeviceMatrix, get_context  # noqa: E402


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    phylo, haps, mix = load_workload(frags, 2)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    ctx = get_context()
    t0 = time.perf_counter()
    _, _, dmat, ms = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=False, keep_device=True)
    print("build kernel %.3f ms (call %.1f ms)" % (ms, 1e3 * (time.perf_counter() - t0)))
    n, h = dmat.shape
    for pinned in (True, False):
        if pinned:
            host = torch.empty((n, h), dtype=torch.float64, pin_memory=True).numpy()
        else:
            host = np.empty((n, h))
        dmat.to_host(out=host)
        args = argparse.Namespace(verbose=False, init_alpha=1.0, tolerance=1e-4,
                                  max_iter=max_iter, n_multi=1)
        for rep in range(2):
            np.random.seed(2)
            t0 = time.perf_counter()
            props, read_mix = mixemt_b200.run_em(host, mix.weights.astype(np.float64), args)
            dt = time.perf_counter() - t0
            print("run_em(pinned=%s) rep %d: %.3f s" % (pinned, rep, dt), flush=True)
            del read_mix
        t0 = time.perf_counter()
        d = DeviceMatrix.from_host(ctx, host)
        print("  upload alone %.1f ms" % (1e3 * (time.perf_counter() - t0)))
        t0 = time.perf_counter()
        out = np.empty((n, h))
        t1 = time.perf_counter()
        d.to_host(out=out)
        t2 = time.perf_counter()
        d.to_host(out=out)
        t3 = time.perf_counter()
        print("  np.empty %.1f ms, download into fresh pageable %.1f ms, into touched %.1f ms"
              % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
        d.free()
        del host, out


if __name__ == "__main__":
    main()
