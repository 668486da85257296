"""Multi-GPU parity check, launched by torchrun (see test_multigpu.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py

rows mode:     every rank runs run_em on its row shard, column sums exchanged per
               iteration (peer stores in the tail kernel, or NCCL); proportions must
               equal the single-GPU run on the full matrix, the shards of the read
               matrix must equal its rows; on a Build-17 matrix (class tiles) the
               sharded run must equal the CPU oracle on the whole matrix.
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
    lo0, hi0 = lo, hi
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

    # rows mode on the layout the bench's multi-GPU numbers run on: a Build-17 matrix of
    # string-sorted signatures goes over class tiles on every rank, the column sums through
    # the peer mailboxes of the tail kernel; checked against the CPU oracle on the whole
    # matrix, not against another run of the same library
    from mixemt_b200 import synth
    from mixemt_b200.phylo_tables import PhyloTables
    from mixemt_b200.preprocess import HapVarBaseMatrix, SignatureCSR, build_matrix_from_csr
    from mixemt_b200._lib import lib, check, ptr
    from oracle import oracle_c
    import ctypes
    phylo = PhyloTables.load(os.path.join(ROOT, "tests", "golden", "phylotree17.npz"))
    haps17 = sorted(phylo.hap_var)
    mix17 = synth.make_mixture(phylo, phylo.refseq, [("H1", 0.5), ("L3e", 0.3), ("U5a1", 0.2)],
                               20000, seed=7, strings=False)
    tables = HapVarBaseMatrix(phylo.refseq, phylo, haps17).pack()
    csr = mix17.csr(tables)
    n17 = csr.n_rows
    lo, hi = sharding.row_shard(n17, rank, world)
    a_, b_ = int(csr.row_ptr[lo]), int(csr.row_ptr[hi])
    sub = SignatureCSR(csr.row_ptr[lo:hi + 1] - csr.row_ptr[lo], csr.pos_idx[a_:b_],
                       csr.base_code[a_:b_])
    _, _, shard17, _ = build_matrix_from_csr(tables, sub, ctx=ctx, want_host=False,
                                             keep_device=True)
    w17 = mix17.weights.astype(np.float64)
    sess = ctypes.c_void_p()
    check(lib.mxb_em_create(ctx.handle, shard17.handle, ptr(w17[lo:hi].copy()), 1,
                            ctypes.byref(sess)))
    nb, flag = ctypes.c_int64(), ctypes.c_int64()
    check(lib.mxb_em_pass_bytes(sess, ctypes.byref(nb), ctypes.byref(flag)))
    lib.mxb_em_destroy(sess)
    assert flag.value == 0, "the shard was expected to run over class tiles"
    init17 = np.log(np.random.RandomState(8).dirichlet([1.0] * len(haps17), size=1))
    a17 = make_args(max_iter=150, tolerance=1e-7, n_multi=1, b200_shard="rows")
    p17, m17, info17, _ = em.run_em_device(shard17, w17[lo:hi], a17, inits=init17)
    full17, _ = oracle_c.build_matrix(tables, csr, want_counts=False)
    o_p, o_m, o_it = oracle_c.run_em(full17, w17, init17, a17.max_iter, a17.tolerance)
    assert list(o_it) == info17["iterations"], (o_it, info17)
    err_p = np.abs(p17 - o_p).max()
    fin = np.isfinite(o_m[lo:hi])
    err_m = np.abs(m17[fin] - o_m[lo:hi][fin]).max()
    assert err_p < 1e-10 and err_m < 1e-9, (err_p, err_m)
    err_p17 = err_p
    assert np.array_equal(np.argmax(m17, 1), np.argmax(o_m[lo:hi], 1))
    shard17.free()

    # default start (inits=None): every rank has its own numpy.random state, rank 0's draws
    # must reach all ranks (em.py:36 consumes the global stream)
    np.random.seed(1000 + rank)
    p_def, _, info_def, _ = em.run_em_device(shard, wts[lo0:hi0], make_args(
        max_iter=25, tolerance=1e-9, n_multi=1, b200_shard="rows"), want_host=False)
    np.random.seed(1000)
    want_init = np.log(np.random.dirichlet([1.0] * h)).reshape(1, h)
    p_chk, _, _, _ = em.run_em_device(full, wts, make_args(max_iter=25, tolerance=1e-9),
                                      inits=want_init, want_host=False)
    assert np.abs(p_def - p_chk).max() < 1e-12, np.abs(p_def - p_chk).max()

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
        print("MULTIGPU_OK world=%d rows:iters=%s tiles-vs-oracle:iters=%s dprops=%.2e restarts ok "
              "exchange=%s" % (world, info_rows["iterations"], info17["iterations"], err_p17, how))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
