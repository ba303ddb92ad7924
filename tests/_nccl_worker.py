"""Worker of tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU, NCCL): the
row-sharded apply with the halo exchange overlapped with the interior tile rows, checked on hardware
against the oracle's apply of the GLOBAL problem (1e-12), and against the non-overlapped form."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    from spuq_b200.sharding import RowShardedApply, row_ranges
    from tests.cases import fd_case
    dev = _abi.device()
    n, M = 64, 4
    idx = synthetic.anisotropic_indices(M, 7.0)
    L = idx.shape[0]
    c = fd_case(n, M, idx, "hermite", 0.5)                         # global problem + oracle result (host)
    y0, y1 = row_ranges(n, world, align=1)[rank]
    nyl = y1 - y0 + 2
    rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, y0, y1, device=dev)
    sh = RowShardedApply(DeviceBlocks(rowptr, colidx, vals, (nyl * n, nyl * n)), c.rvs, n, y0, y1)
    Wg = torch.from_numpy(np.ascontiguousarray(c.W.T)).view(L, n, n)
    W = torch.full((L, nyl, n), 7.7e77, dtype=torch.float64)        # poison the halo rows: they must be overwritten
    W[:, 1:-1, :] = Wg[:, y0:y1, :]
    if rank == 0:
        W[:, 0, :] = 0.0                                             # outside the grid: Dirichlet boundary
    if rank == world - 1:
        W[:, -1, :] = 0.0
    W = W.view(L, nyl * n).to(dev)
    w = sp.SlabMultiVector(idx, slab=W, _sorted=True)
    plan = sh.plan_for(w)
    errs = []
    results = []
    for k, (halo, overlap) in enumerate((("peer", False), ("nccl", False), ("nccl", True), ("peer", False), ("peer", False))):
        sh.halo = halo
        # the owner fills a new slab (halo rows included) LATE on one rank: the neighbour's rows must not arrive
        # before that fill and be overwritten by it (the ready-to-receive handshake of k_halo_push)
        if k >= 3 and rank == k % world:
            torch.cuda._sleep(40_000_000)
        Wc = W.clone()
        Y = torch.full_like(Wc, float("nan"))
        sh.apply(plan, Wc, Y, overlap=overlap)
        torch.cuda.synchronize()
        Yown = Y.view(L, nyl, n)[:, 1:-1, :].cpu().numpy().reshape(L, -1)
        ref = c.Y.T.reshape(L, n, n)[:, y0:y1, :].reshape(L, -1)
        errs.append(float(np.max(np.linalg.norm(Yown - ref, axis=1) / np.linalg.norm(c.Y.T, axis=1))))
        results.append(Yown)
    same = max(float(np.max(np.abs(results[0] - r))) for r in results[1:])
    # repeated exchanges on one buffer (the flag protocol must not let a writer overtake the reader)
    sh.halo = "peer"
    Wc = W.clone()
    Y = torch.empty_like(Wc)
    for _ in range(25):
        sh.apply(plan, Wc, Y)
    torch.cuda.synchronize()
    same = max(same, float(np.max(np.abs(Y.view(L, nyl, n)[:, 1:-1, :].cpu().numpy().reshape(L, -1) - results[0]))))
    t = torch.tensor([max(errs), same], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("NCCL_PARITY kernel=%s relerr=%.3e overlap_vs_plain=%.3e launches=%d" % (plan.kernel_name, float(t[0]), float(t[1]), sh.last_launches))
        assert float(t[0]) <= 1e-12 and float(t[1]) == 0.0
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
