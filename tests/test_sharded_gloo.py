"""N > 1 path on the CPU: two gloo ranks, each with its M-shard of the operator on the host-emulated
library, combined along Lambda; the gathered result must match the oracle (1e-12)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import conftest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spuq_b200 as sp
        from spuq_b200 import _abi, lambda_tables, synthetic
        from spuq_b200.multi_operator import DeviceBlocks
        from spuq_b200.sharding import ShardedApply
        from tests.cases import fd_case
        _abi.use_library(conftest.EMU_LIB, "cpu")
        idx = synthetic.anisotropic_indices(6, 8.0)
        c = fd_case(8, 6, idx, "hermite", 0.5)
        L = idx.shape[0]
        tables = lambda_tables.LambdaTables(idx, c.rvs)
        blocks = DeviceBlocks(c.rowptr, c.colidx, c.vals, (c.N, c.N))
        sh = ShardedApply(blocks, c.rvs, tables)
        assert sh.world == world and sh.ranges[0][0] == 0 and sh.ranges[-1][1] == 6
        W = torch.from_numpy(np.ascontiguousarray(c.W.T))          # L x N slab
        w = sp.SlabMultiVector(idx, slab=W, _sorted=True)
        plan = sh.plan_for(w)
        Yfull = torch.zeros((sh.cols * world, c.N), dtype=torch.float64)
        Yown = torch.empty((sh.cols, c.N), dtype=torch.float64)
        sh.apply(plan, W, Yfull, Yown)
        parts = [torch.empty_like(Yown) for _ in range(world)]
        dist.all_gather(parts, Yown)
        if rank == 0:
            Y = torch.cat(parts)[:L].numpy().T
            err = float(np.max(np.linalg.norm(Y - c.Y, axis=0)) / np.linalg.norm(c.Y))
            with open(out_path, "w") as f:
                f.write("%r\n" % err)
    finally:
        dist.destroy_process_group()


def test_m_sharded_apply_two_ranks(tmp_path, emu):
    out = tmp_path / "err.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert float(out.read_text()) <= 1e-12


def _row_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spuq_b200 as sp
        from spuq_b200 import _abi, synthetic
        from spuq_b200.multi_operator import DeviceBlocks
        from spuq_b200.sharding import RowShardedApply, row_ranges
        from tests.cases import fd_case
        _abi.use_library(conftest.EMU_LIB, "cpu")
        n, M = 16, 4
        idx = synthetic.complete_order_indices(M, 2)
        c = fd_case(n, M, idx, "hermite", 0.5)                     # global problem + oracle result
        L = idx.shape[0]
        y0, y1 = row_ranges(n, world)[rank]
        nyl = y1 - y0 + 2
        rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, y0, y1)
        blocks = DeviceBlocks(rowptr, colidx, vals, (nyl * n, nyl * n))
        sh = RowShardedApply(blocks, c.rvs, n, y0, y1)
        Wg = torch.from_numpy(np.ascontiguousarray(c.W.T)).view(L, n, n)      # global slab, L x ny x nx
        W = torch.zeros((L, nyl, n), dtype=torch.float64)
        W[:, 1:-1, :] = Wg[:, y0:y1, :]                            # owned rows only: halos come from the exchange
        W = W.view(L, nyl * n)
        w = sp.SlabMultiVector(idx, slab=W, _sorted=True)
        plan = sh.plan_for(w)
        Y = torch.empty_like(W)
        sh.apply(plan, W, Y)
        own = Y.view(L, nyl, n)[:, 1:-1, :].contiguous()
        parts = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(parts, own)                                # equal shares: n = 16, world = 2
        if rank == 0:
            Yg = torch.cat(parts, dim=1).reshape(L, n * n).numpy().T
            err = float(np.max(np.linalg.norm(Yg - c.Y, axis=0)) / np.linalg.norm(c.Y))
            with open(out_path, "w") as f:
                f.write("%r\n" % err)
    finally:
        dist.destroy_process_group()


def test_row_sharded_apply_two_ranks(tmp_path, emu):
    out = tmp_path / "err.txt"
    mp.spawn(_row_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert float(out.read_text()) <= 1e-12


def _pcg_worker(rank, world, port, out_path, precond):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spuq_b200 as sp
        from spuq_b200 import _abi, synthetic
        from spuq_b200.multi_operator import DeviceBlocks
        from spuq_b200.sharding import RowShardedApply, RowShardedPCG, row_ranges
        from tests.cases import fd_case
        from oracle import linalg as olin, sg as osg
        _abi.use_library(conftest.EMU_LIB, "cpu")
        n, M = 16, 3
        idx = synthetic.complete_order_indices(M, 2)
        c = fd_case(n, M, idx, "legendre")                         # global problem, x* = c.W
        L = idx.shape[0]
        ranges = row_ranges(n, world)
        y0, y1 = ranges[rank]
        nyl = y1 - y0 + 2
        rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, y0, y1)
        blocks = DeviceBlocks(rowptr, colidx, vals, (nyl * n, nyl * n))
        sh = RowShardedApply(blocks, c.rvs, n, y0, y1)
        Fg = torch.from_numpy(np.ascontiguousarray(c.Y.T)).view(L, n, n)      # f = A x*, global
        F = torch.zeros((L, nyl, n), dtype=torch.float64)
        F[:, 1:-1, :] = Fg[:, y0:y1, :]
        F = F.view(L, nyl * n)
        w = sp.SlabMultiVector(idx, slab=torch.zeros_like(F), _sorted=True)
        plan = sh.plan_for(w)
        d = float(c.mats[0].diagonal()[0])
        solver = RowShardedPCG(sh, n, n, ranges, mean_stencil=(d, -1.0, -1.0) if precond.startswith("mean") else None,
                               mean_solver={"mean-transpose": "transpose", "mean-partition": "partition"}.get(precond, "strip"))
        X, zeta, it = solver.solve(plan, F, torch.zeros_like(F), eps=1e-9, maxiter=200)
        own = X.view(L, nyl, n)[:, 1:-1, :].contiguous()
        parts = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(parts, own)
        if rank == 0:
            Xg = torch.cat(parts, dim=1).reshape(L, n * n).numpy().T
            # the oracle's PCG on the same numbers, with the same preconditioner
            basis = olin.CanonicalBasis(c.N)
            if precond.startswith("mean"):
                lu = olin.ScipySolveOperator(c.mats[0], basis, basis, factor=True)
                P = osg.PreconditioningOperator("mean", lambda b, f: lu)
            else:
                class _Id(olin.Operator):
                    def apply(self, v):
                        return v.copy()
                P = _Id()
            oh = []
            x_o, zeta_o, it_o = osg.pcg(c.A, c.A * c.w, P, 0 * c.w, eps=1e-9, maxiter=200, history=oh)
            k = min(len(oh), len(solver.last_history), 6)
            traj = float(np.max(np.abs(np.array(solver.last_history[:k]) - np.array(oh[:k])) / np.array(oh[:k])))
            err = float(np.max(np.abs(Xg - x_o.to_slab())))
            with open(out_path, "w") as f:
                f.write("%r %r %r %r\n" % (err, traj, it, it_o))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precond", ["mean", "mean-partition", "mean-transpose", "identity"])
def test_row_sharded_pcg_two_ranks(tmp_path, emu, precond):
    """PCG on row-sharded slabs (halo exchange for A, scalars all-reduced in the device state, mean
    preconditioner by the strip solver with a gathered separator system, by the tridiagonal partition method
    or through a transposition to column shards) against the oracle's PCG on the global problem."""
    out = tmp_path / "res.txt"
    mp.spawn(_pcg_worker, args=(2, _free_port(), str(out), precond), nprocs=2, join=True)
    err, traj, it, it_o = [float(v) for v in out.read_text().split()]
    assert abs(it - it_o) <= 1, (it, it_o)
    assert traj <= 1e-9 and err <= 1e-8


def _strip_rows_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sps
        import scipy.sparse.linalg as spla
        from spuq_b200 import _abi
        from spuq_b200.sharding import StripRowsMeanSolver
        _abi.use_library(conftest.EMU_LIB, "cpu")
        worst = 0.0
        for nx, ny, bounds, L in [(40, 75, [0, 30, 52, 75], 3), (33, 120, [0, 41, 80, 120], 2), (12, 9, [0, 3, 6, 9], 2)]:
            ranges = list(zip(bounds[:-1], bounds[1:]))
            T = lambda n, o, dd: sps.diags([o, dd, o], [-1, 0, 1], shape=(n, n))      # noqa: E731
            A0 = sps.csr_matrix(sps.kron(sps.identity(ny), T(nx, -1.0, 2.0)) + sps.kron(T(ny, -1.0, 2.0), sps.identity(nx)))
            rng = np.random.RandomState(17)
            R = rng.random_sample((L, ny, nx))
            r0, r1 = ranges[rank]
            Rc = torch.from_numpy(np.ascontiguousarray(R[:, r0:r1, :]).reshape(L, -1))
            lu = spla.splu(A0.tocsc())
            ref = np.stack([lu.solve(R[k].reshape(-1)) for k in range(L)]).reshape(L, ny, nx)[:, r0:r1, :].reshape(L, -1)
            for schur in ("sharded", "gather"):                    # separator system by column shards / redundantly
                solver = StripRowsMeanSolver(nx, ranges, rank, (4.0, -1.0, -1.0), schur=schur)
                Sc = torch.empty_like(Rc)
                for _ in range(2):                                 # second call: buffers reused
                    solver.solve(Rc, Sc)
                worst = max(worst, float(np.max(np.abs(Sc.numpy() - ref)) / np.max(np.abs(ref))))
                # the band where it lies inside halo-padded slabs (no copy), with <R, S> from the final pass
                n_loc = r1 - r0
                Rp = torch.full((L, n_loc + 2, nx), 7.0, dtype=torch.float64)
                Sp = torch.full((L, n_loc + 2, nx), -3.0, dtype=torch.float64)
                Rp[:, 1:-1, :] = Rc.view(L, n_loc, nx)
                d = torch.zeros(1, dtype=torch.float64)
                solver.solve(Rp[:, 1:-1, :], Sp[:, 1:-1, :], dot_out=d)
                worst = max(worst, float(np.max(np.abs(Sp[:, 1:-1, :].reshape(L, -1).numpy() - ref)) / np.max(np.abs(ref))))
                assert float(Sp[:, 0, :].min()) == -3.0 == float(Sp[:, -1, :].max())       # halo rows untouched
                dref = float(np.sum(Rc.numpy() * ref))
                worst = max(worst, abs(float(d[0]) - dref) / abs(dref))
        t = torch.tensor([worst], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            with open(out_path, "w") as f:
                f.write("%r\n" % float(t[0]))
    finally:
        dist.destroy_process_group()


def test_strip_rows_mean_solver_three_ranks(tmp_path, emu):
    """The strip solver on row bands of three ranks (uneven bands, several strips per band, x-chunks with
    separators, a band of three rows) against the global sparse direct solve."""
    out = tmp_path / "err.txt"
    mp.spawn(_strip_rows_worker, args=(3, _free_port(), str(out)), nprocs=3, join=True)
    assert float(out.read_text()) <= 1e-12
