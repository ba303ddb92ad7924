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
