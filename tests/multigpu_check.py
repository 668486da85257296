"""Multi-GPU parity check, launched by torchrun (see test_multigpu.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py

rows mode:     every rank runs run_em on its row shard with the per-iteration
               NCCL all-reduce; proportions must equal the single-GPU run on the
               full matrix, the shards of the read matrix must equal its rows.
restarts mode: restarts dealt over the ranks and combined must equal the
               single-GPU multi-restart run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import make_args
    from mixemt_b200 import em, sharding
    from mixemt_b200.runtime import DeviceMatrix, get_context

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = get_context()
    ctx.init_comm_from_torch()

    rs = np.random.RandomState(5)
    n, h = 4001, 5408
    mat = -rs.gamma(2.0, 8.0, size=(n, h))
    mat[rs.rand(n, h) < 0.4] = 0.0
    wts = rs.randint(1, 30, size=n)
    inits = np.log(rs.dirichlet([1.0] * h, size=5))

    # single-GPU truth (same on every rank, no communication)
    full = DeviceMatrix.from_host(ctx, mat)
    a1 = make_args(max_iter=120, tolerance=1e-7, n_multi=1)
    p_one, m_one, info_one, _ = em.run_em_device(full, wts, a1, inits=inits[:1])

    # rows mode
    lo, hi = sharding.row_shard(n, rank, world)
    shard = DeviceMatrix.from_host(ctx, mat[lo:hi])
    a_rows = make_args(max_iter=120, tolerance=1e-7, n_multi=1, b200_shard="rows")
    p_rows, m_rows, info_rows, _ = em.run_em_device(shard, wts[lo:hi], a_rows, inits=inits[:1])
    assert info_rows["iterations"] == info_one["iterations"], (info_rows, info_one)
    err_p = np.abs(p_rows - p_one).max()
    err_m = np.abs(m_rows - m_one[lo:hi]).max()
    assert err_p < 1e-12 and err_m < 1e-9, (err_p, err_m)
    gathered = [None] * world
    dist.all_gather_object(gathered, p_rows.tobytes())
    assert all(g == gathered[0] for g in gathered), "ranks disagree on the proportions"

    # rows mode with an empty shard: fewer signature rows than ranks
    tiny = mat[:world - 1]
    lo, hi = sharding.row_shard(len(tiny), rank, world)
    a_t = make_args(max_iter=6, tolerance=1e-12, n_multi=1, b200_shard="rows")
    p_t, m_t, info_t, _ = em.run_em_device(DeviceMatrix.from_host(ctx, tiny[lo:hi]),
                                           wts[:world - 1][lo:hi], a_t, inits=inits[:1])
    p_t1, _, info_t1, _ = em.run_em_device(DeviceMatrix.from_host(ctx, tiny), wts[:world - 1],
                                           make_args(max_iter=6, tolerance=1e-12), inits=inits[:1])
    assert info_t["iterations"] == info_t1["iterations"] and m_t.shape == (hi - lo, h)
    assert np.abs(p_t - p_t1).max() < 1e-12, np.abs(p_t - p_t1).max()

    # restarts mode
    a5 = make_args(max_iter=60, tolerance=1e-7, n_multi=5)
    p_all, m_all, _, _ = em.run_em_device(full, wts, a5, inits=inits)
    a5r = make_args(max_iter=60, tolerance=1e-7, n_multi=5, b200_shard="restarts")
    p_fan, m_fan, info_fan, _ = em.run_em_device(full, wts, a5r, inits=inits)
    assert info_fan["restarts"] == sharding.restart_shard(5, rank, world)
    err_p = np.abs(p_fan - p_all).max()
    fin = np.isfinite(m_all)
    err_m = np.abs(m_fan[fin] - m_all[fin]).max()
    assert err_p < 1e-13 and err_m < 1e-9, (err_p, err_m)

    dist.barrier()
    if rank == 0:
        from mixemt_b200._lib import lib
        how = "p2p" if lib.mxb_comm_p2p_enabled(ctx.handle) else "nccl"
        print("MULTIGPU_OK world=%d rows:iters=%s restarts ok exchange=%s"
              % (world, info_rows["iterations"], how))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
