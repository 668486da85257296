"""ctypes wrapper of oracle/_build/liboracle.so (see oracle.c).

TEST INFRASTRUCTURE ONLY: checker and CPU baseline, never the product path.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_run_em.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def expected_table(tables):
    """Dense H x P expected-symbol table from a packed mixemt_b200
    HapVarBaseMatrix (markers over the reference row)."""
    exp = np.tile(tables.ref_code, (tables.n_hap, 1))
    hap_of = np.repeat(np.arange(tables.n_hap), np.diff(tables.marker_ptr))
    exp[hap_of, tables.marker_pos_idx] = tables.marker_code
    return np.ascontiguousarray(exp)


def build_matrix(tables, csr, want_counts=True, threads=None):
    lib = load()
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    exp = expected_table(tables)
    n = csr.n_rows
    out = np.empty((n, tables.n_hap))
    cnt = np.empty((n, tables.n_hap), dtype=np.int32) if want_counts else None
    lib.orc_build_matrix(ctypes.c_int64(n), ctypes.c_int32(tables.n_hap),
                         ctypes.c_int32(tables.n_pos), _p(exp), _p(tables.hit), _p(tables.miss),
                         _p(csr.row_ptr), _p(csr.pos_idx), _p(np.ascontiguousarray(csr.base_code)),
                         _p(out), _p(cnt))
    return out, cnt


def em_step(mat, weights, ln_props):
    lib = load()
    mat = np.ascontiguousarray(mat, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    lp = np.ascontiguousarray(ln_props, dtype=np.float64)
    n, h = mat.shape
    z = np.empty_like(mat)
    new = np.empty(h)
    lib.orc_em_step(_p(mat), _p(w), _p(lp), ctypes.c_int64(n), ctypes.c_int64(h), _p(z), _p(new))
    return z, new


def run_em(mat, weights, init_lnprops, max_iter, tol):
    lib = load()
    mat = np.ascontiguousarray(mat, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    init = np.ascontiguousarray(init_lnprops, dtype=np.float64).reshape(-1, mat.shape[1])
    n, h = mat.shape
    props = np.empty(h)
    mix = np.empty_like(mat)
    iters = np.zeros(len(init), dtype=np.int64)
    rc = lib.orc_run_em(_p(mat), _p(w), ctypes.c_int64(n), ctypes.c_int64(h), _p(init),
                        ctypes.c_int32(len(init)), ctypes.c_int64(max_iter), ctypes.c_double(tol),
                        _p(props), _p(mix), _p(iters))
    if rc != 0:
        raise MemoryError("oracle run_em")
    return props, mix, iters.tolist()
