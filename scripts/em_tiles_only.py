"""Kernel 2 alone over class tiles on the config-2 workload (rows in the reference's order):
fixed number of EM iterations (ncu target).

    python scripts/em_tiles_only.py [fragments] [iterations]
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import load_workload  # noqa: E402
from mixemt_b200._lib import lib, check, ptr  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import get_context  # noqa: E402


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    phylo, haps, mix = load_workload(frags, 2, strings=True)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    ctx = get_context()
    _, _, dmat, _ = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=False, keep_device=True)
    n, h = dmat.shape
    w = mix.weights.astype(np.float64)
    sess = ctypes.c_void_p()
    check(lib.mxb_em_create(ctx.handle, dmat.handle, ptr(w), 0, ctypes.byref(sess)))
    lnp0 = np.log(np.random.RandomState(1).dirichlet([1.0] * h))
    check(lib.mxb_em_set_lnprops(sess, ptr(lnp0)))
    el, ps = ctypes.c_float(), ctypes.c_float()
    check(lib.mxb_em_iterate_fixed(sess, 5, ctypes.byref(el), None))
    check(lib.mxb_em_iterate_fixed(sess, iters, ctypes.byref(el), ctypes.byref(ps)))
    nb, nd = ctypes.c_int64(), ctypes.c_int64()
    check(lib.mxb_em_pass_bytes(sess, ctypes.byref(nb), ctypes.byref(nd)))
    print("EM %d x %d: %.4f ms per iteration, pass %.4f ms, %.3f GB per pass"
          % (n, h, el.value / iters, ps.value / iters, nb.value / 1e9))
    ms = (ctypes.c_float * 4)()
    check(lib.mxb_em_profile(sess, iters, ms))
    print("per kernel (ms): class sums %.4f, pass %.4f, gather %.4f, tail %.4f" % tuple(ms))
    lib.mxb_em_destroy(sess)


if __name__ == "__main__":
    main()
