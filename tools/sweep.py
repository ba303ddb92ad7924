#!/usr/bin/env python
"""Kernel-variant sweep and memory microbenchmarks on one GPU (development tool, not the bench).

  python tools/sweep.py --config c3 [--variants "1,0,0;2,8,2;2,4,2;..."] [--micro]
Prints one line per variant: kernel name, ms (median of reps, CUDA events), GDoF/s, achieved
algorithmic GB/s and the fraction of the measured HBM peak."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


def micro():
    """Read bandwidth by working-set size with the library's own dot kernel (x . x reads n doubles):
    small sets live in L2, large ones stream from HBM."""
    from spuq_b200 import _abi
    lib = _abi.lib()
    dev = _abi.device()
    scratch = torch.zeros(int(lib.spuq_sg_reduce_scratch_bytes()) // 8 + 1, dtype=torch.float64, device=dev)
    out = torch.zeros(2, dtype=torch.float64, device=dev)
    for mb in (8, 16, 32, 48, 64, 96, 128, 256, 1024, 4096):
        n = mb * (1 << 20) // 8
        x = torch.ones(n, dtype=torch.float64, device=dev)
        fn = lambda: _abi.check(lib.spuq_sg_dot(n, _abi.ptr(x), _abi.ptr(x), _abi.ptr(out), _abi.ptr(scratch), _abi.stream_ptr()))  # noqa: E731
        med, best = timeit(fn, reps=20, warm=5)
        print("micro dot(x,x) %5d MB: %.4f ms  %.0f GB/s (best %.0f)" % (mb, med, n * 8 / med / 1e6, n * 8 / best / 1e6), flush=True)
        del x
    # copy bandwidth of torch (the MEASURED_PEAKS definition) and of our axpy (3 streams)
    n = (2 << 30) // 8
    a = torch.ones(n, dtype=torch.float64, device=dev); b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a), reps=10)
    print("micro torch copy 2 GiB: %.3f ms  %.0f GB/s (best %.0f)" % (med, 2 * n * 8 / med / 1e6, 2 * n * 8 / best / 1e6), flush=True)
    fn = lambda: _abi.check(lib.spuq_sg_axpy(n, 0.5, None, None, _abi.ptr(a), _abi.ptr(b), _abi.stream_ptr()))  # noqa: E731
    med, best = timeit(fn, reps=10)
    print("micro axpy 2 GiB: %.3f ms  %.0f GB/s (best %.0f)" % (med, 3 * n * 8 / med / 1e6, 3 * n * 8 / best / 1e6), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--variants", default="1,0,0;2,8,2;2,4,2;2,16,2;2,8,1;2,16,1;2,4,4")
    ap.add_argument("--paths", default="")
    ap.add_argument("--micro", action="store_true")
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    if os.environ.get("SPUQ_LIB"):                      # A/B builds of the library (development only)
        _abi.use_library(os.environ["SPUQ_LIB"], "cuda:0")
    dev = _abi.device()
    if args.micro:
        micro()
    cfg = bench.CONFIGS[args.config]
    n, M = cfg["n"], cfg["M"]
    N = n * n
    idx = bench.index_set(cfg)
    L = idx.shape[0]
    rvs = synthetic.make_rvs(cfg["kind"], M, cfg["mu"])
    rowptr, colidx, vals = synthetic.fd_blocks(n, M, device=dev)
    blocks = DeviceBlocks(rowptr, colidx, vals, (N, N))
    g = torch.Generator(device=dev); g.manual_seed(1234)
    W = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    Y = torch.empty_like(W)
    w = sp.SlabMultiVector(idx, slab=W, _sorted=True)
    peak, _ = bench.measured_peaks()
    ref = None
    todo = [(_abi.PATH_AUTO, tuple(int(v) for v in t.split(","))) for t in args.variants.split(";") if t]
    for pth in [int(v) for v in args.paths.split(",") if v]:
        todo.append((pth, (0, 0, 0)))
    ops = {}
    for pth, tune in todo:
        if pth not in ops:
            ops[pth] = sp.MultiOperator.from_blocks(blocks, rvs, path=pth)
        plan = ops[pth].plan_for(w)
        info = plan.info()
        try:
            plan.tune(tune[0], tune[1], tune[2], 0)
            med, best = timeit(lambda: plan.apply(W, Y), reps=args.reps)
        except Exception as e:           # keep sweeping
            print("variant %s path %d: FAILED %s" % (tune, pth, e), flush=True)
            continue
        if ref is None:
            ref = Y.clone()
            err = 0.0
        else:
            err = float(torch.linalg.norm(Y - ref) / torch.linalg.norm(ref))
        gbs = info["algorithmic_bytes"] / med / 1e6
        print("variant %-10s path %d %-32s %8.3f ms (best %8.3f)  %7.1f GDoF/s  %6.0f GB/s alg  frac %.3f  "
              "fp64 %.1f TF  diff %.1e" % (tune, pth, plan.kernel_name, med, best, N * L / med / 1e6, gbs, gbs / peak,
                                          info["algorithmic_flops"] / med / 1e9, err), flush=True)
    print(json.dumps({"config": args.config, "N": N, "L": L, "M": M, "info": info}))


if __name__ == "__main__":
    main()
