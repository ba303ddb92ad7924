#!/usr/bin/env python
"""PCG solve on one GPU (BASELINE.json config 4 style): mean-based preconditioner, eps = 1e-10.
  python tools/pcg_bench.py --config c4 [--eps 1e-10] [--maxiter 200]
Prints iterations, time to solution, time per pass and GDoF/s per pass."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--eps", type=float, default=1e-10)
    ap.add_argument("--maxiter", type=int, default=200)
    ap.add_argument("--poll", type=int, default=8)
    ap.add_argument("--mean-solver", default="strip", choices=["strip", "partition", "transpose"])
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_sharded(args)
    torch.cuda.set_device(0)
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    dev = _abi.device()
    cfg = bench.CONFIGS[args.config]
    n, M = cfg["n"], cfg["M"]
    N = n * n
    idx = bench.index_set(cfg)
    L = idx.shape[0]
    rvs = synthetic.make_rvs(cfg["kind"], M, cfg["mu"])
    rowptr, colidx, vals = synthetic.fd_blocks(n, M, device=dev)
    blocks = DeviceBlocks(rowptr, colidx, vals, (N, N))
    A = sp.MultiOperator.from_blocks(blocks, rvs)
    g = torch.Generator(device=dev); g.manual_seed(1234)
    X = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    x_true = sp.SlabMultiVector(idx, slab=X, _sorted=True)
    f = A.apply(x_true)
    # mean block as a host CSR matrix for the preconditioner set-up (constant-coefficient 5-point)
    A0 = synthetic.blocks_to_scipy(rowptr, colidx, vals[:1], N)[0]
    P = sp.PreconditioningOperator("mean", lambda b, f_: sp.ScipySolveOperator(A0, b, b))
    w0 = 0 * f
    hist = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, zeta, it = sp.pcg(A, f, P, w0, eps=args.eps, maxiter=args.maxiter, poll_every=args.poll, history=hist)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = float((x.slab - X).abs().max())

    def ms(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    t_apply = ms(lambda: A.apply(x))
    t_prec = ms(lambda: P * f)
    t_dot = ms(lambda: sp.inner(x, f))
    t_axpy = ms(lambda: x.__iadd__(f))
    passes = max(it - 1, 1)
    print(json.dumps({"config": args.config, "N": N, "L": L, "M": M, "eps": args.eps, "numit": it, "zeta": zeta,
                      "seconds": dt, "ms_per_pass": dt / passes * 1e3, "gdof_per_s_per_pass": N * L * passes / dt / 1e9,
                      "max_abs_error": err, "ms_apply": t_apply, "ms_precond": t_prec, "ms_dot": t_dot, "ms_axpy": t_axpy, "apply_kernel": A.plan_for(x_true).kernel_name,
                      "zeta_history": hist[:6] + ["..."] + hist[-3:]}))


def main_sharded(args):
    """torchrun --nproc-per-node G tools/pcg_bench.py ...: the same solve with the grid rows sharded over
    G GPUs (spuq_b200.sharding.RowShardedPCG): halo exchange for A, all-reduced scalars, mean
    preconditioner through an all-to-all transposition to column shards."""
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    from spuq_b200.sharding import RowShardedApply, RowShardedPCG, row_ranges
    _abi.reset()
    dev = _abi.device()
    cfg = bench.CONFIGS[args.config]
    n, M = cfg["n"], cfg["M"]
    idx = bench.index_set(cfg)
    L = idx.shape[0]
    rvs = synthetic.make_rvs(cfg["kind"], M, cfg["mu"])
    ranges = row_ranges(n, world, align=1)
    y0, y1 = ranges[rank]
    nyl = y1 - y0 + 2
    rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, y0, y1, device=dev)
    blocks = DeviceBlocks(rowptr, colidx, vals, (nyl * n, nyl * n))
    sh = RowShardedApply(blocks, rvs, n, y0, y1)
    # manufactured solution: x* ~ U(0,1) on the owned rows (seeded per rank), f = A x*
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    X = torch.zeros((L, nyl, n), dtype=torch.float64, device=dev)
    X[:, 1:-1, :] = torch.rand((L, y1 - y0, n), dtype=torch.float64, device=dev, generator=g)
    X = X.view(L, nyl * n)
    w = sp.SlabMultiVector(idx, slab=X, _sorted=True)
    plan = sh.plan_for(w)
    F = torch.empty_like(X)
    sh.apply(plan, X.clone(), F)
    solver = RowShardedPCG(sh, n, n, ranges, mean_stencil=(4.0, -1.0, -1.0), mean_solver=args.mean_solver)
    solver._zero_halos(F)
    tmp = torch.empty_like(F)
    solver.precondition(F, tmp)                        # warm-up: solver set-up, cuBLAS handle, NCCL channels
    sh.apply(plan, tmp, torch.empty_like(F))
    del tmp
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    Xs, zeta, it = solver.solve(plan, F, torch.zeros_like(F), eps=args.eps, maxiter=args.maxiter)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    err = (Xs.view(L, nyl, n)[:, 1:-1, :] - X.view(L, nyl, n)[:, 1:-1, :]).abs().max().reshape(1)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    passes = max(it - 1, 1)
    if rank == 0:
        print(json.dumps({"config": args.config, "n_gpus": world, "sharding": "grid rows; mean preconditioner: " + solver.mean_solver,
                          "N": n * n, "L": L, "M": M, "eps": args.eps, "numit": it, "zeta": zeta, "seconds": dt,
                          "ms_per_pass": dt / passes * 1e3, "gdof_per_s_per_pass": n * n * L * passes / dt / 1e9,
                          "max_abs_error": float(err[0]), "zeta_history": solver.last_history[:6] + ["..."] + solver.last_history[-3:]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
