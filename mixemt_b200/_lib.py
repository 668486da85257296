"""ctypes binding of libmixemt_b200.so (the C-ABI declared in include/mixemt_b200.h).

There is no CPU fallback: if the shared library has not been built the import
fails loudly, and every compute entry point raises when no CUDA device is
visible.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmixemt_b200.so")

MXB_OK, MXB_ERR_CUDA, MXB_ERR_ARG, MXB_ERR_VALUE, MXB_ERR_KEY, MXB_ERR_NOMEM, \
    MXB_ERR_RANGE = range(7)
MXB_EM_SHARDED = 1
MXB_EM_RAW = 2

c_void_pp = ctypes.POINTER(ctypes.c_void_p)
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_dbl = ctypes.c_double
P = ctypes.c_void_p  # every buffer / handle crosses the ABI as a plain pointer


class LibraryMissing(ImportError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise LibraryMissing(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mixemt_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    # Let the dlopen("libnccl.so.2") inside the library find torch's bundled NCCL.
    if "MXB_NCCL_LIB" not in os.environ:
        try:
            import nvidia.nccl  # type: ignore
            for base in nvidia.nccl.__path__:
                cand = os.path.join(base, "lib", "libnccl.so.2")
                if os.path.isfile(cand):
                    os.environ["MXB_NCCL_LIB"] = cand
                    break
        except Exception:  # pragma: no cover - NCCL is optional for 1 GPU
            pass
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()

# name -> (restype, argtypes); mirrors include/mixemt_b200.h one to one.
_PROTOS = {
    "mxb_abi_version": (ctypes.c_int, []),
    "mxb_last_error": (ctypes.c_char_p, []),
    "mxb_device_count": (ctypes.c_int, []),
    "mxb_ctx_create": (ctypes.c_int, [ctypes.c_int, c_void_pp]),
    "mxb_ctx_destroy": (ctypes.c_int, [P]),
    "mxb_ctx_set_stream": (ctypes.c_int, [P, P]),
    "mxb_ctx_synchronize": (ctypes.c_int, [P]),
    "mxb_ctx_trim": (ctypes.c_int, [P]),
    "mxb_ctx_launch_count": (c_i64, [P]),
    "mxb_comm_unique_id": (ctypes.c_int, [P]),
    "mxb_comm_init": (ctypes.c_int, [P, P, ctypes.c_int, ctypes.c_int]),
    "mxb_comm_destroy": (ctypes.c_int, [P]),
    "mxb_comm_p2p_enabled": (ctypes.c_int, [P]),
    "mxb_comm_allreduce_host": (ctypes.c_int, [P, P, c_i64, ctypes.c_int]),
    "mxb_sig_count": (ctypes.c_int, [P, P, c_i64, P]),
    "mxb_sig_parse": (ctypes.c_int, [P, P, c_i64, P, c_i64, P, P, P, P,
                                     ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "mxb_reduce_reads": (ctypes.c_int, [P, P, P, c_i64, c_void_pp]),
    "mxb_sigset_sizes": (ctypes.c_int, [P, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                        ctypes.POINTER(c_i64)]),
    "mxb_sigset_export": (ctypes.c_int, [P, P, P, P, P, P, P, P, P, P]),
    "mxb_sigset_destroy": (ctypes.c_int, [P]),
    "mxb_phylo_pack": (ctypes.c_int, [P, c_i32, c_i32, c_i32, P, P, P, P, P, P, c_void_pp]),
    "mxb_phylo_destroy": (ctypes.c_int, [P]),
    "mxb_build_matrix": (ctypes.c_int, [P, P, c_i64, P, P, P, P, P, c_void_pp,
                                        ctypes.POINTER(ctypes.c_float)]),
    "mxb_matrix_alloc": (ctypes.c_int, [P, c_i64, c_i64, c_void_pp]),
    "mxb_matrix_upload": (ctypes.c_int, [P, P, c_i64, c_i64, c_void_pp]),
    "mxb_matrix_download": (ctypes.c_int, [P, P, P]),
    "mxb_matrix_shape": (ctypes.c_int, [P, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "mxb_matrix_data": (P, [P]),
    "mxb_matrix_argmax_rows": (ctypes.c_int, [P, P, P]),
    "mxb_matrix_destroy": (ctypes.c_int, [P]),
    "mxb_matrix_gather_cols": (ctypes.c_int, [P, P, P, c_i64, c_void_pp]),
    "mxb_matrix_vote_count": (ctypes.c_int, [P, P, P, P, P]),
    "mxb_assign_reads": (ctypes.c_int, [P, P, P, P, c_i32, c_dbl, P]),
    "mxb_matrix_download_rows": (ctypes.c_int, [P, P, c_i64, c_i64, P]),
    "mxb_matrix_upload_rows": (ctypes.c_int, [P, P, c_i64, c_i64, P]),
    "mxb_matrix_fold_ranks": (ctypes.c_int, [P, P, c_dbl]),
    "mxb_em_create": (ctypes.c_int, [P, P, P, ctypes.c_int, c_void_pp]),
    "mxb_em_destroy": (ctypes.c_int, [P]),
    "mxb_em_set_lnprops": (ctypes.c_int, [P, P]),
    "mxb_em_iterate": (ctypes.c_int, [P, c_i64, c_dbl, ctypes.POINTER(c_i64),
                                      ctypes.POINTER(c_i32)]),
    "mxb_em_pass_bytes": (ctypes.c_int, [P, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "mxb_em_iterate_fixed": (ctypes.c_int, [P, c_i64, ctypes.POINTER(ctypes.c_float),
                                            ctypes.POINTER(ctypes.c_float)]),
    "mxb_em_get_lnprops": (ctypes.c_int, [P, ctypes.c_int, P]),
    "mxb_em_read_mix": (ctypes.c_int, [P, P, ctypes.c_int, c_dbl]),
    "mxb_run_em": (ctypes.c_int, [P, P, P, c_i64, c_i64, P, c_i32, c_i64, c_dbl, c_i32,
                                  P, P, P, P]),
    "mxb_run_em_dev": (ctypes.c_int, [P, P, P, P, c_i32, c_i64, c_dbl, c_i32, P, P,
                                      c_void_pp, P, P]),
    "mxb_em_step": (ctypes.c_int, [P, P, P, P, c_i64, c_i64, P, P]),
}

EXPORTED_SYMBOLS = tuple(sorted(_PROTOS))

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class MixemtB200Error(RuntimeError):
    """CUDA / NCCL / argument failure inside the native library."""


class NumericRangeError(MixemtB200Error, ArithmeticError):
    pass


def last_error():
    msg = lib.mxb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    """Turn a non-zero status into the exception the reference would raise
    (ValueError / KeyError) or a MixemtB200Error."""
    if rc == MXB_OK:
        return
    msg = last_error() or what
    if rc == MXB_ERR_VALUE:
        raise ValueError(msg)
    if rc == MXB_ERR_KEY:
        raise KeyError(msg)
    if rc == MXB_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == MXB_ERR_RANGE:
        raise NumericRangeError(msg)
    raise MixemtB200Error("%s (status %d)" % (msg, rc))


def ptr(arr):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return ctypes.c_void_p(arr.ctypes.data)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
