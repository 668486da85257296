"""Kernel 2 on the config-2 workload in each row layout: class tiles (default), fp64 rows
(MXB_EM_NO_PACK=1).  Fixed iterations for timing, then a run to convergence: iteration counts
must agree and the log-proportions must agree to rounding.

    python scripts/em_modes.py [fragments] [iterations]
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import load_workload  # noqa: E402
from mixemt_b200._lib import lib, check, ptr  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import get_context  # noqa: E402

MODES = [("class tiles", {}), ("fp64 rows", {"MXB_EM_NO_PACK": "1"})]


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    phylo, haps, mix = load_workload(frags, 2, strings=True)   # rows in the reference order
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    csr = mix.csr(tables)
    ctx = get_context()
    _, _, dmat, _ = build_matrix_from_csr(tables, csr, ctx=ctx, want_host=False, keep_device=True)
    n, h = dmat.shape
    w = mix.weights.astype(np.float64)
    lnp0 = np.log(np.random.RandomState(1).dirichlet([1.0] * h))
    ref = None
    for name, env in MODES:
        os.environ.pop("MXB_EM_NO_PACK", None)
        os.environ.update(env)
        sess = ctypes.c_void_p()
        ctx.synchronize()
        t0 = time.perf_counter()
        check(lib.mxb_em_create(ctx.handle, dmat.handle, ptr(w), 0, ctypes.byref(sess)))
        ctx.synchronize()
        t_create = time.perf_counter() - t0
        check(lib.mxb_em_set_lnprops(sess, ptr(lnp0)))
        el, ps = ctypes.c_float(), ctypes.c_float()
        check(lib.mxb_em_iterate_fixed(sess, 5, ctypes.byref(el), None))
        check(lib.mxb_em_iterate_fixed(sess, iters, ctypes.byref(el), ctypes.byref(ps)))
        nb, nd = ctypes.c_int64(), ctypes.c_int64()
        check(lib.mxb_em_pass_bytes(sess, ctypes.byref(nb), ctypes.byref(nd)))
        lnp_fixed = np.empty(h)
        check(lib.mxb_em_get_lnprops(sess, 0, ptr(lnp_fixed)))
        # to convergence from the same start
        check(lib.mxb_em_set_lnprops(sess, ptr(lnp0)))
        it, conv = ctypes.c_int64(), ctypes.c_int32()
        t0 = time.perf_counter()
        check(lib.mxb_em_iterate(sess, 10000, 1e-4, ctypes.byref(it), ctypes.byref(conv)))
        t_conv = time.perf_counter() - t0
        lnp_conv = np.empty(h)
        check(lib.mxb_em_get_lnprops(sess, 0, ptr(lnp_conv)))
        lib.mxb_em_destroy(sess)
        if ref is None:
            ref = (lnp_fixed, lnp_conv, it.value)
        live = np.isfinite(ref[0]) & np.isfinite(lnp_fixed)
        d_fixed = float(np.abs(lnp_fixed[live] - ref[0][live]).max())
        d_props = float(np.abs(np.exp(lnp_conv) - np.exp(ref[1])).max())
        print("%-16s %d x %d: create %.1f ms, %.4f ms per iteration (pass %.4f ms), %.3f GB per "
              "pass; converged=%d after %d iterations in %.3f s; max |d ln pi| after %d fixed "
              "iterations %.3g, max |d pi| at convergence %.3g (vs %s)"
              % (name, n, h, t_create * 1e3, el.value / iters, ps.value / iters, nb.value / 1e9,
                 conv.value, it.value, t_conv, iters + 5, d_fixed, d_props, MODES[0][0]))


if __name__ == "__main__":
    main()
