"""Host logic of the class-tile pass (csrc/tile_plan.h through mxb_tile_plan; no GPU): the plan
covers every row of every batch exactly once with contiguous ranges per CTA, segments are cut at
copy boundaries, copies never split the rows a warp handles in one step, the thread -> row
mapping of tile_pass_kernel (em_tiles.cuh) visits every row of a segment exactly once, and the
CTAs' modelled times are balanced.  The reference has no counterpart (em.py:57-91 is three numpy
expressions over the whole matrix)."""
import ctypes
import math

import numpy as np
import pytest

from mixemt_b200._lib import lib, check, last_error, MXB_ERR_ARG

ROWS = 128
STEP_CYCLES = [440, 440, 440, 440, 440, 600, 1000, 1300, 1626, 1900]


def plan(n_cls, n_rows, num_sms=148):
    n_cls = np.ascontiguousarray(n_cls, dtype=np.int32)
    cap = 8 * len(n_cls) + 4 * num_sms + 16
    seg = np.zeros((cap, 8), dtype=np.int32)
    out = [ctypes.c_int32() for _ in range(5)]
    check(lib.mxb_tile_plan(ctypes.c_void_p(n_cls.ctypes.data), len(n_cls), n_rows, num_sms,
                            ctypes.c_void_p(seg.ctypes.data), cap, *[ctypes.byref(o) for o in out]))
    n_cta, n_seg, n_slots, slot_bytes, errors = [o.value for o in out]
    assert n_seg <= cap
    return seg[:n_seg], n_cta, n_slots, slot_bytes, errors


def lg_of(n_cls):
    lg = 3
    while ((n_cls + 1) >> 1) > (8 << lg):
        lg += 1
    return lg


def check_plan(n_cls, n_rows, num_sms=148):
    seg, n_cta, n_slots, slot_bytes, errors = plan(n_cls, n_rows, num_sms)
    assert errors == 0
    assert 1 <= n_cta <= min(num_sms, max(1, math.ceil(n_rows / 32)))
    assert 2 <= n_slots <= 6 and n_slots * slot_bytes <= 160 * 1024
    nb = len(n_cls)
    # CTAs in order, every CTA a contiguous range of rows of the whole matrix
    assert (np.diff(seg[:, 0]) >= 0).all() and seg[0, 0] == 0 and seg[-1, 0] == n_cta - 1
    assert set(seg[:, 0].tolist()) == set(range(n_cta))
    pos = 0
    for cta, b, r0, rows, fit, copies, lg, r_pad in seg.tolist():
        assert b * ROWS + r0 == pos, "gap or overlap in the row ranges"
        pos += rows
        batch_rows = min(ROWS, n_rows - b * ROWS)
        assert rows >= 1 and r0 + rows <= batch_rows
        assert lg == lg_of(int(n_cls[b])) and 3 <= lg <= 9
        assert r_pad == (int(n_cls[b]) + 1 + 3) // 4 * 4
        per_step = 32 >> lg if lg < 5 else 1
        assert fit % per_step == 0 and fit >= per_step
        assert fit * r_pad * 8 <= slot_bytes, "a copy must fit into a ring slot"
        assert copies == math.ceil(rows / fit)
        # what the kernel's bulk copies and its exchange scratch rely on: 16-byte aligned
        # copies of at most 2^20 - 1 bytes, and the row slices' shares fit the 64 KB scratch
        assert (r_pad * 8) % 32 == 0 and fit * r_pad * 8 < 2 ** 20
        chunks = (int(n_cls[b]) + 1) >> 1
        n_slices = (512 >> lg) if lg >= 5 else 16
        assert n_slices * chunks * 16 <= 64 * 1024
        assert chunks <= 8 << lg                 # eight double2 chunks per thread
        # a segment ends at a copy boundary or at the end of its batch (or takes the crumbs)
        assert rows % fit == 0 or r0 + rows == batch_rows
    assert pos == n_rows
    return seg, n_cta, slot_bytes


def visit_counts(rows, fit, lg):
    """Rows of a segment as the threads of tile_pass_kernel visit them (512 threads, thread t
    handles rows (t >> lg) + (512 >> lg) i, copy by copy)."""
    count = np.zeros(rows, dtype=int)
    sweep = 512 >> lg
    copies = math.ceil(rows / fit)
    for first in range(min(sweep, rows + sweep)):       # one representative thread per row slot
        row = first
        warp_row = first if lg >= 5 else (first >> (5 - lg)) << (5 - lg)
        for c in range(copies):
            row0, row_end = c * fit, min(c * fit + fit, rows)
            wrow = warp_row + (row - first)
            while wrow < row_end:
                assert wrow >= row0, "a warp step straddles two copies"
                if row < row_end:
                    count[row] += 1
                row += sweep
                wrow += sweep
    return count


def modelled_cycles(seg, slot_bytes):
    warp, mem = {}, {}
    for cta, b, r0, rows, fit, copies, lg, r_pad in seg.tolist():
        w = 4200.0
        for c in range(copies):
            n = min(fit, rows - c * fit)
            w += 500.0 + math.ceil(n / (512 >> lg)) * STEP_CYCLES[lg]
        warp[cta] = warp.get(cta, 0.0) + w
        mem[cta] = mem.get(cta, 0.0) + rows * r_pad * 8 / 22.0
    return np.array([max(warp[c], mem[c]) for c in sorted(warp)])


@pytest.mark.parametrize("seed", range(6))
def test_random_plans_cover_every_row_once(seed):
    rs = np.random.RandomState(seed)
    nb = int(rs.randint(1, 60))
    n_rows = nb * ROWS - int(rs.randint(0, ROWS))
    n_cls = rs.choice([1, 3, 60, 63, 64, 65, 130, 255, 256, 300, 511, 512, 513, 700, 1144, 2047,
                       2048, 2668, 4096, 5408, 8191, 8192], size=nb)
    seg, n_cta, slot_bytes = check_plan(n_cls, n_rows, num_sms=int(rs.choice([4, 37, 148])))
    for cta, b, r0, rows, fit, copies, lg, r_pad in seg.tolist():
        assert (visit_counts(rows, fit, lg) == 1).all(), (rows, fit, lg)


def test_config2_like_plan_is_balanced():
    """1083 batches with the class counts of the config-2 matrix (median 135, mean 235, a tail
    up to 2256): 148 CTAs whose modelled times differ by a few per cent."""
    rs = np.random.RandomState(0)
    n_cls = np.minimum(2256, np.maximum(20, rs.lognormal(np.log(135), 0.9, size=1083))).astype(int)
    n_rows = 138569
    seg, n_cta, slot_bytes = check_plan(n_cls, n_rows)
    assert n_cta == 148
    t = modelled_cycles(seg, slot_bytes)
    assert t.max() < 1.08 * t.mean(), (t.max(), t.mean())
    # a batch is split over few CTAs: at most one extra segment per CTA boundary
    assert len(seg) <= 1083 + 148


@pytest.mark.parametrize("median,min_bytes", [(135, 0), (500, 2 ** 32)])
def test_config3_shard_plan(median, min_bytes):
    """A config-3 shard (1.25 M rows = 9766 batches per GPU, BASELINE.json configs[2]; measured
    tiles: ~1.8 GB) and the same rows with four times as many classes per batch (tiles beyond
    2^32 bytes: offsets are 64-bit): the plan covers every row once with 148 balanced CTAs."""
    rs = np.random.RandomState(3)
    n_rows = 1250000
    nb = (n_rows + ROWS - 1) // ROWS
    n_cls = np.minimum(5408, np.maximum(20, rs.lognormal(np.log(median), 0.9, size=nb))).astype(int)
    seg, n_cta, slot_bytes = check_plan(n_cls, n_rows)
    assert n_cta == 148 and len(seg) <= nb + 148
    tile_bytes = sum(rows * r_pad * 8 for _, _, _, rows, _, _, _, r_pad in seg.tolist())
    assert tile_bytes > min_bytes
    t = modelled_cycles(seg, slot_bytes)
    assert t.max() < 1.02 * t.mean(), (t.max(), t.mean())


def test_tiny_and_extreme_shapes():
    check_plan([1], 1)
    check_plan([8192], 128)                  # 64 KB rows: two slots of one row
    check_plan([8192, 1], 129)
    seg, n_cta, _ = check_plan([100] * 3, 300)
    assert n_cta <= 10                        # ceil(300 / 32)
    seg, n_cta, _ = check_plan([5408] * 4, 512, num_sms=148)
    assert n_cta == 16


def test_bad_arguments_are_rejected():
    n_cls = np.array([10, 10], dtype=np.int32)
    p = ctypes.c_void_p(n_cls.ctypes.data)
    assert lib.mxb_tile_plan(p, 2, 128, 148, None, 0, None, None, None, None, None) == MXB_ERR_ARG
    assert lib.mxb_tile_plan(p, 2, 300, 148, None, 0, None, None, None, None, None) == MXB_ERR_ARG
    bad = np.array([0], dtype=np.int32)
    assert lib.mxb_tile_plan(ctypes.c_void_p(bad.ctypes.data), 1, 5, 148, None, 0, None, None,
                             None, None, None) == MXB_ERR_ARG
    assert "mxb_tile_plan" in last_error()
