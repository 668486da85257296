"""Host-side rules of the multi-GPU modes.  Pure numpy; the
collectives are passed in as callables so that the same code runs over NCCL
(``Context.allreduce_host`` / device kernels) and, in the CPU tests, over a
world_size-2 gloo group.

* rows mode:     signature rows are split into contiguous balanced shards; each
                 EM iteration all-reduces the H partial column sums T_j.
* restarts mode: the matrix is replicated, the restarts are dealt in contiguous
                 blocks, and the per-rank results are combined as the reference
                 combines restarts (em.py:145-163): log-proportions are summed
                 and divided by n_multi, read matrices are folded with
                 logaddexp and shifted by -log(n_multi).
"""
import numpy as np


def row_shard(n_rows, rank, world):
    """Contiguous balanced row range [lo, hi) of ``rank``."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def restart_shard(n_multi, rank, world):
    """Indices of the restarts ``rank`` runs: contiguous balanced blocks, so that folding the
    ranks' partial results in rank order adds the restarts in restart order (em.py:155-156)."""
    return list(range((n_multi * rank) // world, (n_multi * (rank + 1)) // world))


def combine_restart_props(local_sum_lnprops, n_multi, allreduce):
    """em.py:155 + :158-163 across ranks: ``allreduce(arr, "sum")`` in place."""
    total = allreduce(np.ascontiguousarray(local_sum_lnprops, dtype=np.float64), "sum")
    if n_multi > 1:
        total = total / n_multi
    return np.exp(total)


def fold_read_mix(local_mix, n_multi, allreduce):
    """log(sum over ranks of exp(local_mix)) - log(n_multi), computed as
    max-shift + sum so that it needs only max/sum all-reduces; ``local_mix`` is
    this rank's logaddexp fold of its restarts (all -inf when it ran none).  A
    collective-only statement of what ``mxb_matrix_fold_ranks`` computes on the
    GPUs (there: all-to-all of row shards + logaddexp in rank order), used by
    the gloo tests."""
    mx = allreduce(np.array(local_mix, dtype=np.float64, copy=True), "max")
    with np.errstate(invalid="ignore"):
        lin = np.where(np.isinf(mx), np.where(mx < 0, 0.0, 1.0), np.exp(local_mix - mx))
    lin = allreduce(np.ascontiguousarray(lin), "sum")
    with np.errstate(divide="ignore"):
        out = np.where(np.isinf(mx), mx, mx + np.log(lin))
    return out - (np.log(n_multi) if n_multi > 1 else 0.0)


def share_from_rank0(make, shape, rank, allreduce):
    """Rank 0 evaluates ``make()`` (e.g. the Dirichlet draws of em.py:36 from its global
    numpy.random stream); every rank returns that array.  Shared as a sum all-reduce of
    zeros on the other ranks -- ``-inf`` entries (a zero proportion) survive a sum, a NaN
    would not and is rejected."""
    if rank == 0:
        arr = np.ascontiguousarray(make(), dtype=np.float64).reshape(shape)
        if np.isnan(arr).any():
            raise ValueError("initial log-proportions hold NaN")
    else:
        arr = np.zeros(shape, dtype=np.float64)
    return allreduce(arr, "sum")
