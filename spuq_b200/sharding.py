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

__all__ = ["shard_ranges", "padded_columns", "ShardedApply", "row_ranges", "RowShardedApply"]


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


# ---- row sharding ---------------------------------------------------------------------------------
def row_ranges(ny, world, align=8):
    """Contiguous grid-row ranges per rank, multiples of `align` rows (the kernel's tile height)
    except for the last one."""
    per = -(-ny // world)
    per = -(-per // align) * align
    return [(min(r * per, ny), min((r + 1) * per, ny)) for r in range(world)]


class RowShardedApply:
    """The SG operator with the SPATIAL rows sharded over the ranks (axis 3).

    Rank r owns the grid rows [y0, y1) for all |Lambda| columns; its local slab carries one halo grid
    row above and below (local grid (y1 - y0 + 2) x nx).  One apply = exchange of the two boundary
    rows with the neighbouring ranks (2 x nx x L doubles per neighbour, point to point) followed by
    the single-GPU kernel on the local grid; output halo rows are not meaningful.  No reduction of
    Y is needed, so the traffic per rank is 1/world of the single-GPU traffic plus the halos.

    ``blocks``: local blocks from synthetic.fd_blocks_rows(n, M, y0, y1) (or any assembly of the same
    rows); ``W`` slabs are L x ((y1-y0+2)*nx)."""

    def __init__(self, blocks, rvs, nx, y0, y1, group=None, path=0):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nx, self.y0, self.y1 = nx, y0, y1
        self.nyl = y1 - y0 + 2
        self.op = MultiOperator.from_blocks(blocks, rvs, path=path)
        self._bufs = None

    def plan_for(self, w):
        return self.op.plan_for(w)

    def exchange_halos(self, W):
        """Fill the halo rows of the local slab W (L x nyl*nx) from the neighbours' boundary rows."""
        if self.world == 1:
            return
        L = W.shape[0]
        W3 = W.view(L, self.nyl, self.nx)
        if self._bufs is None or self._bufs[0].shape[0] != L:
            mk = lambda: torch.empty((L, self.nx), dtype=W.dtype, device=W.device)      # noqa: E731
            self._bufs = (mk(), mk(), mk(), mk())
        s_up, s_dn, r_up, r_dn = self._bufs
        ops = []
        if self.rank > 0:
            s_up.copy_(W3[:, 1, :])                              # my first owned row -> bottom halo of rank-1
            ops += [dist.P2POp(dist.isend, s_up, self.rank - 1, self.group),
                    dist.P2POp(dist.irecv, r_up, self.rank - 1, self.group)]
        if self.rank < self.world - 1:
            s_dn.copy_(W3[:, self.nyl - 2, :])                   # my last owned row -> top halo of rank+1
            ops += [dist.P2POp(dist.isend, s_dn, self.rank + 1, self.group),
                    dist.P2POp(dist.irecv, r_dn, self.rank + 1, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if self.rank > 0:
            W3[:, 0, :].copy_(r_up)
        if self.rank < self.world - 1:
            W3[:, self.nyl - 1, :].copy_(r_dn)

    def apply(self, plan, W, Y):
        self.exchange_halos(W)
        plan.apply(W, Y)
        return Y
