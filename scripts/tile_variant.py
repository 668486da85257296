"""One library variant over class tiles on the config-2 workload: per-kernel times, a checksum of
the proportions after a fixed number of iterations, and (trace builds, -DMXB_TILE_TRACE) the
per-team timeline of one pass.

    MXB_VARIANT_LIB=scripts/variants/lib_x.so python scripts/tile_variant.py [fragments] [iters] [tag]

Development tool: the variant library is copied over mixemt_b200/lib/libmixemt_b200.so of the
(scratch) GPU-box snapshot before the package is imported.
"""
import ctypes
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

variant = os.environ.get("MXB_VARIANT_LIB")
if variant:
    shutil.copyfile(os.path.join(ROOT, variant), os.path.join(ROOT, "mixemt_b200", "lib", "libmixemt_b200.so"))

from bench import load_workload  # noqa: E402
from mixemt_b200._lib import lib, check, ptr  # noqa: E402
from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr  # noqa: E402
from mixemt_b200.runtime import get_context  # noqa: E402


def main():
    frags = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    tag = sys.argv[3] if len(sys.argv) > 3 else "v"
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    os.environ["MXB_TILE_TRACE_ITEMS"] = os.path.join(out_dir, "trace_items_%s.txt" % tag)
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
    best = 1e9
    for _ in range(3):
        check(lib.mxb_em_iterate_fixed(sess, iters, ctypes.byref(el), ctypes.byref(ps)))
        best = min(best, el.value / iters)
    nb, nd = ctypes.c_int64(), ctypes.c_int64()
    check(lib.mxb_em_pass_bytes(sess, ctypes.byref(nb), ctypes.byref(nd)))
    lnp = np.empty(h)
    check(lib.mxb_em_get_lnprops(sess, 1, ptr(lnp)))
    p = np.exp(lnp)
    top = np.argsort(-p)[:3]
    ms = (ctypes.c_float * 4)()
    check(lib.mxb_em_profile(sess, iters, ms))
    print("[%s] %d x %d: %.4f ms/iter (best of 3), %.3f GB/pass | pi %.4f pass %.4f gather %.4f tail %.4f | "
          "props %s sum %.15f"
          % (tag, n, h, best, nb.value / 1e9, ms[0], ms[1], ms[2], ms[3],
             " ".join("%s=%.12f" % (haps[i], p[i]) for i in top), p.sum()), flush=True)
    if hasattr(lib, "mxb_debug_tile_trace"):
        check(lib.mxb_em_iterate_fixed(sess, 1, ctypes.byref(el), None))
        words = 160 * 16 * 64 * 2
        buf = np.zeros(words, dtype=np.uint64)
        lib.mxb_debug_tile_trace.restype = ctypes.c_int
        rc = lib.mxb_debug_tile_trace(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(words))
        print("trace rc", rc)
        np.save(os.path.join(out_dir, "trace_%s.npy" % tag), buf.reshape(160, 16, 64, 2))
    lib.mxb_em_destroy(sess)


if __name__ == "__main__":
    main()
