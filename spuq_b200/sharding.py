"""Multi-GPU form of the SG operator application (BASELINE.json north_star): the M expansion terms
are sharded over the ranks, every rank applies its terms to a full replica of w, and the partial
results are summed with a reduce-scatter along the Lambda (column) axis, so that rank r ends up
owning the r-th contiguous block of output columns.

The reference has no multi-process code at all; the split is the sum over m
of application/egsz/multi_operator2.py:59-82, which is independent per m.
"""
import numpy as np
import torch
import torch.distributed as dist

from .multi_operator import MultiOperator

__all__ = ["shard_ranges", "padded_columns", "ShardedApply"]


def shard_ranges(tables, M, world):
    """Contiguous m-ranges per rank, balanced by term count (the L terms of the mean block go to
    rank 0).  Deterministic, so that every rank computes the same split."""
    w = np.array([(tables.nbr_up[m] >= 0).sum() + (tables.nbr_dn[m] >= 0).sum()
                  + (tables.beta_0[m] != 0).sum() for m in range(M)], dtype=np.float64)
    total = w.sum() + tables.L
    target = total / world
    ranges, lo, acc = [], 0, float(tables.L)
    for g in range(world):
        hi = lo
        if g == world - 1:
            hi = M
        else:
            while hi < M and acc + w[hi] / 2 <= target * (g + 1):
                acc += w[hi]
                hi += 1
        ranges.append((lo, hi))
        lo = hi
    return ranges


def padded_columns(L, world):
    """Columns per rank after padding L up to a multiple of the world size."""
    return (L + world - 1) // world


class ShardedApply:
    """y_own = (A w)[columns of this rank] with A = sum over ranks of the rank's term shard.

    ``blocks`` / ``rvs`` as for MultiOperator.from_blocks; ``tables`` the LambdaTables of the full
    index set.  ``apply(W, Yfull, Yown)``: W is the L x N slab (full replica), Yfull an Lp x N work
    slab (Lp = world * padded_columns), Yown receives this rank's Lp/world columns."""

    def __init__(self, blocks, rvs, tables, group=None, path=0):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        M = tables.maxm
        self.ranges = shard_ranges(tables, M, self.world)
        self.op = MultiOperator.from_blocks(blocks, rvs, m_range=self.ranges[self.rank],
                                            include_mean=(self.rank == 0), path=path)
        self.L = tables.L
        self.cols = padded_columns(self.L, self.world)

    def plan_for(self, w):
        return self.op.plan_for(w)

    def apply(self, plan, W, Yfull, Yown):
        assert Yfull.shape[0] == self.cols * self.world and Yown.shape[0] == self.cols
        plan.apply(W, Yfull[:self.L])
        if self.world == 1:
            Yown.copy_(Yfull)
            return Yown
        if dist.get_backend(self.group) == "nccl":
            dist.reduce_scatter_tensor(Yown, Yfull, op=dist.ReduceOp.SUM, group=self.group)
        else:
            # gloo (CPU tests) has no reduce-scatter on tensors: all-reduce and keep the own block
            tmp = Yfull.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            Yown.copy_(tmp[self.rank * self.cols:(self.rank + 1) * self.cols])
        return Yown
