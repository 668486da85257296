"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding with
a per-iteration all-reduce of the column sums, and restart fan-out with the
reference's averaging rules.  The per-shard arithmetic is a numpy restatement
of what csrc/em.cu does on each GPU; the full-matrix oracle is the checker."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mixemt_b200 import sharding
from oracle import oracle_np


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def gloo_allreduce(arr, op):
    t = torch.from_numpy(arr)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return arr


def problem(n=90, h=37, seed=0):
    rs = np.random.RandomState(seed)
    mat = -rs.gamma(2.0, 6.0, size=(n, h))
    mat[rs.rand(n, h) < 0.2] = mat.max()
    wts = rs.randint(1, 9, size=n).astype(float)
    inits = np.log(rs.dirichlet([1.0] * h, size=5))
    return mat, wts, inits


def linear_partial_sums(mat_shard, wts_shard, ln_props):
    """One rank's share of an EM iteration in the linear-space form of
    csrc/em.cu: T_j = sum_i w_i L_ij / (sum_k L_ik pi_k), L = exp(M - rowmax)."""
    lin = np.exp(mat_shard - mat_shard.max(axis=1, keepdims=True))
    s = lin @ np.exp(ln_props)
    return (wts_shard / s) @ lin


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mat, wts, inits = problem()
    try:
        if mode == "rows":
            lo, hi = sharding.row_shard(mat.shape[0], rank, world)
            ln = inits[0].copy()
            for _ in range(25):
                t = gloo_allreduce(linear_partial_sums(mat[lo:hi], wts[lo:hi], ln), "sum")
                s = np.exp(ln) * t
                ln = np.log(s / s.sum())
            out[rank] = ln
        elif mode == "inits":
            # every rank has its own numpy.random state; rank 0's Dirichlet draws must reach all
            np.random.seed(100 + rank)
            h = mat.shape[1]
            got = sharding.share_from_rank0(
                lambda: np.log(np.random.dirichlet([1.0] * h, size=3)), (3, h), rank, gloo_allreduce)
            out[rank] = got
        else:
            mine = sharding.restart_shard(len(inits), rank, world)
            acc = np.zeros(mat.shape[1])
            fold = np.full(mat.shape, -np.inf)
            for i in mine:   # what MXB_EM_RAW returns: plain sums / logaddexp fold
                p, m, _ = oracle_np.run_em(mat, wts, inits[i:i + 1], 300, 1e-6)
                acc += np.log(p)
                fold = np.logaddexp(fold, m)
            props = sharding.combine_restart_props(acc, len(inits), gloo_allreduce)
            mix = sharding.fold_read_mix(fold, len(inits), gloo_allreduce)
            out[rank] = (props, mix)
    finally:
        dist.destroy_process_group()


def run_world(mode, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, free_port(), mode, out), nprocs=world, join=True)
    return dict(out)


def test_row_and_restart_shards_partition():
    for n in (0, 1, 7, 138569):
        for world in (1, 2, 3, 8):
            bounds = [sharding.row_shard(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1
    for n_multi in (1, 3, 100):
        for world in (1, 2, 8):
            got = sorted(sum((sharding.restart_shard(n_multi, r, world) for r in range(world)), []))
            assert got == list(range(n_multi))


def test_rows_mode_two_ranks_matches_full_matrix_oracle():
    res = run_world("rows")
    mat, wts, inits = problem()
    ln = inits[0]
    for _ in range(25):
        _, ln = oracle_np.em_step(mat, wts, ln)
    assert np.array_equal(res[0], res[1])          # every rank holds identical proportions
    assert np.abs(np.exp(res[0]) - np.exp(ln)).max() < 1e-13


def test_restart_mode_two_ranks_matches_reference_averaging():
    res = run_world("restarts")
    mat, wts, inits = problem()
    props, mix, _ = oracle_np.run_em(mat, wts, inits, 300, 1e-6)
    for rank in (0, 1):
        assert np.abs(res[rank][0] - props).max() < 1e-13
        assert np.abs(res[rank][1] - mix).max() < 1e-11
    assert abs(props.sum() - 1.0) > 1e-9        # geometric mean is not renormalised (F4)


def test_fold_handles_ranks_without_restarts():
    ident = lambda arr, op: arr                  # world of one
    x = np.array([[0.0, -np.inf], [-3.0, -700.0]])
    assert np.allclose(sharding.fold_read_mix(x, 1, ident), x)
    assert np.allclose(sharding.fold_read_mix(x, 4, ident)[1], x[1] - np.log(4))


def test_default_inits_are_drawn_on_rank0_and_shared():
    """in both sharding modes every rank must start from the same proportions; the
    reference draws from the global numpy.random stream (em.py:36), here rank 0's."""
    res = run_world("inits")
    np.random.seed(100)
    want = np.log(np.random.dirichlet([1.0] * problem()[0].shape[1], size=3))
    assert np.array_equal(res[0], want) and np.array_equal(res[1], want)


def test_restart_blocks_are_contiguous_and_ordered():
    for n_multi, world in ((100, 8), (5, 2), (3, 8)):
        blocks = [sharding.restart_shard(n_multi, r, world) for r in range(world)]
        flat = sum(blocks, [])
        assert flat == list(range(n_multi))                    # rank order == restart order
        assert max(map(len, blocks)) - min(map(len, blocks)) <= 1
