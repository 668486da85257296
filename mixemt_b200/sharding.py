"""Host-side rules of the multi-GPU modes.  Pure numpy; the
collectives are passed in as callables so that the same code runs over NCCL
(``Context.allreduce_host`` / device kernels) and, in the CPU tests, over a
world_size-2 gloo group.

* rows mode:     signature rows are split into contiguous balanced shards; each
                 EM iteration all-reduces the H partial column sums T_j.
* restarts mode: the matrix is replicated, restart i runs on rank i % world,
                 and the per-rank results are combined as the reference
                 combines restarts (em.py:145-163): log-proportions are summed
                 and divided by n_multi, read matrices are folded with
                 logaddexp and shifted by -log(n_multi).
"""
import numpy as np


def row_shard(n_rows, rank, world):
    """Contiguous balanced row range [lo, hi) of ``rank``."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def restart_shard(n_multi, rank, world):
    """Indices of the restarts ``rank`` runs (round robin)."""
    return list(range(rank, n_multi, world))


def combine_restart_props(local_sum_lnprops, n_multi, allreduce):
    """em.py:155 + :158-163 across ranks: ``allreduce(arr, "sum")`` in place."""
    total = allreduce(np.ascontiguousarray(local_sum_lnprops, dtype=np.float64), "sum")
    if n_multi > 1:
        total = total / n_multi
    return np.exp(total)


def fold_read_mix(local_mix, n_multi, allreduce):
    """log(sum over ranks of exp(local_mix)) - log(n_multi), computed as
    max-shift + sum so that it needs only max/sum all-reduces (NCCL has no
    logaddexp); ``local_mix`` is this rank's logaddexp fold of its restarts
    (all -inf when it ran none).  Same arithmetic as the device kernels
    fold_exp_kernel / fold_log_kernel in csrc/em.cu."""
    mx = allreduce(np.array(local_mix, dtype=np.float64, copy=True), "max")
    with np.errstate(invalid="ignore"):
        lin = np.where(np.isinf(mx), np.where(mx < 0, 0.0, 1.0), np.exp(local_mix - mx))
    lin = allreduce(np.ascontiguousarray(lin), "sum")
    with np.errstate(divide="ignore"):
        out = np.where(np.isinf(mx), mx, mx + np.log(lin))
    return out - (np.log(n_multi) if n_multi > 1 else 0.0)
