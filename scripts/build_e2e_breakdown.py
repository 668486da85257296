"""Stage timing of the drop-in build_em_matrix on the config-2 signature strings."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import load_workload  # noqa: E402
from mixemt_b200 import preprocess  # noqa: E402
from mixemt_b200._lib import lib, ptr  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, _flatten_signatures, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import get_context  # noqa: E402


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    phylo, haps, mix = load_workload(frags, 2, strings=True)
    reads = list(mix.signatures)
    ctx = get_context()
    for rep in range(2):
        t = [time.perf_counter()]
        hv = HapVarBaseMatrix(phylo.refseq, phylo, haplogroups=[])
        t.append(time.perf_counter())
        hv.haplogroups = haps
        hv.pack()
        t.append(time.perf_counter())
        buf, offsets = _flatten_signatures(reads)
        t.append(time.perf_counter())
        n = len(reads)
        row_ptr = np.zeros(n + 1, dtype=np.int64)
        lib.mxb_sig_count(buf, ptr(offsets), n, ptr(row_ptr))
        t.append(time.perf_counter())
        csr, err = preprocess.parse_signatures(reads, hv)
        t.append(time.perf_counter())
        hv.to_device(ctx)
        ctx.synchronize()
        t.append(time.perf_counter())
        out, _, _, ms = build_matrix_from_csr(hv, csr, ctx=ctx, want_host=True)
        t.append(time.perf_counter())
        names = ["HapVarBaseMatrix", "pack", "flatten", "sig_count", "parse_signatures (all)",
                 "phylo to device", "kernel + D2H (kernel %.2f ms)" % ms]
        print("rep %d (OMP_NUM_THREADS=%s, cores %d): " % (rep, os.environ.get("OMP_NUM_THREADS"),
                                                        os.cpu_count())
              + ", ".join("%s %.3f" % (nm, b - a) for nm, a, b in zip(names, t[:-1], t[1:])),
              flush=True)
        del out
    t0 = time.perf_counter()
    import argparse
    mat = preprocess.build_em_matrix(phylo.refseq, phylo, reads, haps, argparse.Namespace(verbose=False))
    print("build_em_matrix total %.3f s" % (time.perf_counter() - t0))


if __name__ == "__main__":
    main()
