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
    args = ap.parse_args()
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


if __name__ == "__main__":
    main()
