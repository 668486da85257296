"""Device-side forms of the functions that consume ``run_em``'s results
(SURVEY.md 8f, rows N2-N4).  Same names, arguments and return values as the
reference functions they mirror; the N x H matrices are read where they live,
in HBM, so only O(N) or O(H) numbers cross PCIe:

    reduce_em_matrix          mixemt/preprocess.py:230-251
    find_contribs_from_reads  mixemt/assemble.py:102-124 (_find_contribs_from_reads)
    read_votes                mixemt/stats.py:34-46      (the argmax in report_read_votes)
    assign_read_indexes       mixemt/assemble.py:267-334
    save_npy / load_npy       bin/mixemt:214-245, :168-211 (the .npy halves of -s / -l)

Every matrix argument may be a ``DeviceMatrix``, an ndarray that
``build_em_matrix`` / ``run_em`` handed out in resident mode (its HBM copy is
reused), or any other ndarray (uploaded first: there is no host code path).
"""
import collections
import ctypes

import numpy

from . import _lib
from ._lib import lib, check, ptr
from .runtime import (DeviceMatrix, get_context, lookup_resident, remember_resident,
                      resident_mode)


class _OnDevice(object):
    """Context manager: a DeviceMatrix view of ``mat``, freed on exit if it was
    uploaded just for this call."""

    def __init__(self, mat, ctx=None):
        self.owned = False
        if isinstance(mat, DeviceMatrix):
            self.dev = mat
            return
        dev = lookup_resident(mat) if isinstance(mat, numpy.ndarray) else None
        if dev is None:
            arr = _lib.as_f64(mat)
            if arr.ndim != 2:
                raise ValueError("expected a 2-dimensional matrix")
            dev = DeviceMatrix.from_host(ctx or get_context(), arr)
            self.owned = True
        self.dev = dev

    def __enter__(self):
        return self.dev

    def __exit__(self, *exc):
        if self.owned:
            self.dev.free()


def gather_columns(dev_mat, indexes):
    """``dev_mat[:, indexes]`` as a new ``DeviceMatrix``."""
    cols = numpy.ascontiguousarray(indexes, dtype=numpy.int64)
    handle = ctypes.c_void_p()
    check(lib.mxb_matrix_gather_cols(dev_mat.ctx.handle, dev_mat.handle, ptr(cols), len(cols),
                                     ctypes.byref(handle)))
    return DeviceMatrix(dev_mat.ctx, handle)


def reduce_em_matrix(em_mat, haplogroups, contrib_props):
    """
    Keeps only the columns of the haplogroups listed in ``contrib_props``
    (reference preprocess.py:230-251).  Returns ``(matrix, new_haps)``: a
    ``DeviceMatrix`` when ``em_mat`` is one, otherwise a host ndarray whose HBM
    copy stays registered so that the refinement ``run_em`` (bin/mixemt:318)
    starts without a host->device copy.
    """
    haps_to_keep = {con[1] for con in contrib_props}
    indexes = [i for i in range(len(haplogroups)) if haplogroups[i] in haps_to_keep]
    new_haps = [haplogroups[i] for i in indexes]
    was_resident = isinstance(em_mat, numpy.ndarray) and lookup_resident(em_mat) is not None
    with _OnDevice(em_mat) as dev:
        small = gather_columns(dev, indexes)
    if isinstance(em_mat, DeviceMatrix):
        return small, new_haps
    host = small.to_host()
    if was_resident or resident_mode():
        # opt-in residency only: the HBM copy stays registered behind a read-only array
        host.flags.writeable = False
        remember_resident(host, small)
    else:
        small.free()    # the reference hands back a fresh writable array (preprocess.py:251)
    return host, new_haps


def vote_count(read_hap_mat, wts):
    """``(votes[H], argmax[N])``: weighted votes per haplogroup column from the
    row maxima (first maximum, like ``numpy.argmax(read_hap_mat, 1)``)."""
    with _OnDevice(read_hap_mat) as dev:
        n, h = dev.shape
        weights = numpy.ascontiguousarray(numpy.asarray(wts).reshape(-1), dtype=numpy.int64)
        if weights.shape[0] != n:
            raise ValueError("wts has %d entries for %d rows" % (weights.shape[0], n))
        votes = numpy.zeros(h, dtype=numpy.int64)
        best = numpy.empty(n, dtype=numpy.int64)
        check(lib.mxb_matrix_vote_count(dev.ctx.handle, dev.handle, ptr(weights), ptr(votes),
                                        ptr(best)))
    return votes, best


def find_contribs_from_reads(read_hap_mat, wts, args):
    """
    Column indexes of the haplogroups whose reads (rows voting for them with
    their multiplicity) reach ``args.min_reads`` -- reference
    assemble.py:102-124, in the reference's order (first appearance among the
    row maxima).
    """
    votes, best = vote_count(read_hap_mat, wts)
    seen, first = numpy.unique(best, return_index=True)
    order = seen[numpy.argsort(first, kind="stable")]
    return [int(con) for con in order if votes[con] >= args.min_reads]


def read_votes(read_hap_mat):
    """``numpy.argmax(read_hap_mat, 1)`` of stats.report_read_votes (stats.py:39)."""
    with _OnDevice(read_hap_mat) as dev:
        return dev.argmax_rows()


def assign_rows(read_hap_mat, props, con_indexes, min_fold):
    """int32 per row: position in ``con_indexes`` of the contributor the row is
    assigned to, or -1 (unassigned)."""
    cols = numpy.ascontiguousarray(con_indexes, dtype=numpy.int64)
    log_props = numpy.log(numpy.asarray(props, dtype=numpy.float64))   # assemble.py:300
    con_ln = numpy.ascontiguousarray(log_props[cols])
    with _OnDevice(read_hap_mat) as dev:
        out = numpy.empty(dev.shape[0], dtype=numpy.int32)
        check(lib.mxb_assign_reads(dev.ctx.handle, dev.handle, ptr(cols), ptr(con_ln), len(cols),
                                   float(numpy.log(min_fold)), ptr(out)))
    return out


def assign_read_indexes(contribs, em_results, haps, reads, min_fold):
    """
    Maps contributor names to the set of row indexes assigned to them, plus
    ``'unassigned'`` (reference assemble.py:267-334): a row goes to the
    contributor with the highest ``read_mix - log(props)`` if it leads the next
    contributor by at least ``log(min_fold)``.
    """
    props, read_hap_mat = em_results
    contrib_reads = collections.defaultdict(set)
    if len(contribs) > 1:
        cols = [haps.index(group) for _, group, _ in contribs]
        names = [hap_n for hap_n, _, _ in contribs]
        assigned = assign_rows(read_hap_mat, props, cols, min_fold)[:len(reads)]
        for k, name in enumerate(names):
            rows = numpy.nonzero(assigned == k)[0]
            if len(rows):
                contrib_reads[name].update(rows.tolist())
        rows = numpy.nonzero(assigned < 0)[0]
        if len(rows):
            contrib_reads['unassigned'].update(rows.tolist())
    else:
        contrib_reads[contribs[0][0]].update(range(len(reads)))
    return contrib_reads


# ---- .npy streaming (N4) -------------------------------------------------------
_CHUNK_BYTES = 256 << 20


def save_npy(mat, path):
    """``numpy.save(path, mat)`` for a matrix that lives in HBM: the file is
    written in row chunks, so no N x H host copy is needed (bin/mixemt:240-242
    saves the EM input and result matrices this way)."""
    from numpy.lib import format as npy_format
    with _OnDevice(mat) as dev:
        n, h = dev.shape
        if not str(path).endswith(".npy"):
            path = str(path) + ".npy"          # numpy.save appends the suffix
        rows_per = max(1, _CHUNK_BYTES // max(8 * h, 1))
        buf = numpy.empty((min(rows_per, max(n, 1)), h), dtype=numpy.float64)
        with open(path, "wb") as handle:
            npy_format.write_array_header_1_0(handle, {"descr": "<f8", "fortran_order": False,
                                                       "shape": (n, h)})
            for r0 in range(0, n, rows_per):
                k = min(rows_per, n - r0)
                check(lib.mxb_matrix_download_rows(dev.ctx.handle, dev.handle, r0, k, ptr(buf)))
                buf[:k].tofile(handle)
    return path


def load_npy(path, ctx=None):
    """``numpy.load(path)`` straight into HBM (bin/mixemt:199-203): returns a
    ``DeviceMatrix``; the file is read in row chunks."""
    from numpy.lib import format as npy_format
    ctx = ctx or get_context()
    with open(path, "rb") as handle:
        version = npy_format.read_magic(handle)
        if version == (1, 0):
            shape, fortran, dtype = npy_format.read_array_header_1_0(handle)
        else:
            shape, fortran, dtype = npy_format.read_array_header_2_0(handle)
        if len(shape) != 2 or fortran or dtype != numpy.dtype("<f8"):
            raise ValueError("%s: expected a C-ordered 2-d float64 array, found %s %s%s"
                             % (path, shape, dtype, " (Fortran order)" if fortran else ""))
        n, h = shape
        dev = DeviceMatrix.empty(ctx, n, h)
        rows_per = max(1, _CHUNK_BYTES // max(8 * h, 1))
        for r0 in range(0, n, rows_per):
            k = min(rows_per, n - r0)
            buf = numpy.fromfile(handle, dtype=numpy.float64, count=k * h)
            if buf.size != k * h:
                dev.free()
                raise ValueError("%s: truncated file" % path)
            check(lib.mxb_matrix_upload_rows(ctx.handle, dev.handle, r0, k, ptr(buf)))
    return dev
