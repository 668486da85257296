"""Coded EM pass: every loop variant on the config-2 workload, one process per variant.

    python scripts/em_variants.py [fragments] [iterations]

Variants are selected by environment flags that csrc/em.cu reads once per process
(pick_pass_coded), so each one runs in its own child under `timeout` -- the pipelined
variant synchronises through mbarriers and has not run on a GPU yet; a deadlock must not
take the box down.  Prints ms per iteration, ms per pass and the largest difference of the
log-proportions after the iterations against the default variant (v3 adds the same numbers
in the same order: expected 0; t384 maps columns to threads differently and
`pairs` codes a slightly different set of rows: expected ~1e-15).
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [("default", {}),
            ("v3 pipelined", {"MXB_EM_CODED_V3": "1"}),
            ("v3, 384 threads", {"MXB_EM_CODED_V3": "1", "MXB_EM_CODED_T384": "1"}),
            ("v3, pairs", {"MXB_EM_CODED_V3": "1", "MXB_EM_CODED_PAIRS": "1"}),
            ("v3, pairs, 384 threads", {"MXB_EM_CODED_V3": "1", "MXB_EM_CODED_PAIRS": "1",
                                        "MXB_EM_CODED_T384": "1"}),
            ("v3, pairs, 384, coded only", {"MXB_EM_CODED_V3": "1", "MXB_EM_CODED_PAIRS": "1",
                                            "MXB_EM_CODED_T384": "1", "MXB_EM_CODED_COMPACT": "1"}),
            ("384 threads", {"MXB_EM_CODED_T384": "1"}),
            ("pairs (chunk dictionary)", {"MXB_EM_CODED_PAIRS": "1"}),
            ("pairs, 384 threads", {"MXB_EM_CODED_PAIRS": "1", "MXB_EM_CODED_T384": "1"}),
            ("coded rows only", {"MXB_EM_CODED_COMPACT": "1"}),
            ("pairs, 384, coded only", {"MXB_EM_CODED_PAIRS": "1", "MXB_EM_CODED_T384": "1",
                                        "MXB_EM_CODED_COMPACT": "1"}),
            ("fp64 rows", {"MXB_EM_NO_PACK": "1"})]


def child(frags, iters, out_path):
    from bench import load_workload
    from mixemt_b200._lib import lib, check, ptr
    from mixemt_b200.preprocess import HapVarBaseMatrix, build_matrix_from_csr
    from mixemt_b200.runtime import get_context
    phylo, haps, mix = load_workload(frags, 2)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps).pack()
    ctx = get_context()
    _, _, dmat, _ = build_matrix_from_csr(tables, mix.csr(tables), ctx=ctx, want_host=False,
                                          keep_device=True)
    n, h = dmat.shape
    w = mix.weights.astype(np.float64)
    sess = ctypes.c_void_p()
    check(lib.mxb_em_create(ctx.handle, dmat.handle, ptr(w), 0, ctypes.byref(sess)))
    lnp0 = np.log(np.random.RandomState(1).dirichlet([1.0] * h))
    check(lib.mxb_em_set_lnprops(sess, ptr(lnp0)))
    el, ps = ctypes.c_float(), ctypes.c_float()
    check(lib.mxb_em_iterate_fixed(sess, 5, ctypes.byref(el), None))
    check(lib.mxb_em_iterate_fixed(sess, iters, ctypes.byref(el), ctypes.byref(ps)))
    lnp = np.empty(h)
    check(lib.mxb_em_get_lnprops(sess, 0, ptr(lnp)))
    lib.mxb_em_destroy(sess)
    np.save(out_path, lnp)
    print(json.dumps({"rows": n, "cols": h, "ms_per_iteration": el.value / iters,
                      "ms_per_pass": ps.value / iters}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
        return
    frags = sys.argv[1] if len(sys.argv) > 1 else "1000000"
    iters = sys.argv[2] if len(sys.argv) > 2 else "200"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    ref = None
    for name, env in VARIANTS:
        out = os.path.join(ROOT, "gpurun_out", "variant_%d.npy" % VARIANTS.index((name, env)))
        e = dict(os.environ)
        for k in ("MXB_EM_CODED_V3", "MXB_EM_CODED_T384", "MXB_EM_CODED_PAIRS",
                  "MXB_EM_CODED_COMPACT", "MXB_EM_NO_PACK"):
            e.pop(k, None)
        e.update(env)
        cmd = ["timeout", "120", sys.executable, os.path.abspath(__file__), "--child", frags, iters,
               out]
        r = subprocess.run(cmd, env=e, capture_output=True, text=True)
        if r.returncode != 0:
            print("%-30s FAILED rc=%d %s" % (name, r.returncode, r.stderr.strip()[-300:]))
            continue
        info = json.loads(r.stdout.strip().splitlines()[-1])
        lnp = np.load(out)
        if ref is None:
            ref = lnp
        live = np.isfinite(ref) & np.isfinite(lnp)
        print("%-30s %.4f ms per iteration, %.4f ms per pass, max |d ln pi| vs default %.3g"
              % (name, info["ms_per_iteration"], info["ms_per_pass"],
                 float(np.abs(lnp[live] - ref[live]).max())))


if __name__ == "__main__":
    main()
