#!/usr/bin/env python
"""BASELINE.json config 5: the roofline curve of the batched SG apply -- |Lambda| x N sweep on one GPU.
  python tools/roofline_sweep.py [--out gpurun_out/roofline_sweep.json]
For every point: ms per apply (median of 5 after 3 warm-ups, CUDA events), GDoF/s, algorithmic GB/s and
the fraction of the measured HBM peak.  Legendre chaos, total-degree sets, 5-point FD blocks."""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402

SETS = [(3, 2), (5, 2), (5, 3), (10, 3), (20, 3), (30, 3)]          # |Lambda| = 10, 21, 56, 286, 1771, 5456
GRIDS = [128, 256, 512, 1024, 2048]                                # N = 16k ... 4.2M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/roofline_sweep.json")
    ap.add_argument("--max-gb", type=float, default=100.0)
    ap.add_argument("--budget-s", type=float, default=400.0)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    dev = _abi.device()
    peak, _ = bench.measured_peaks()
    rows, t_start = [], time.time()
    for n in GRIDS:
        for (M, p) in SETS:
            idx = synthetic.complete_order_indices(M, p)
            L, N = idx.shape[0], n * n
            need_gb = (16.0 * N * L + 8.0 * 5 * N * (M + 1) * 2) / 1e9
            if need_gb > args.max_gb or time.time() - t_start > args.budget_s:
                continue
            rvs = synthetic.make_rvs("legendre", M)
            rowptr, colidx, vals = synthetic.fd_blocks(n, M, device=dev)
            A = sp.MultiOperator.from_blocks(DeviceBlocks(rowptr, colidx, vals, (N, N)), rvs)
            W = torch.rand((L, N), dtype=torch.float64, device=dev)
            Y = torch.empty_like(W)
            plan = A.plan_for(sp.SlabMultiVector(idx, slab=W, _sorted=True))
            info = plan.info()
            for _ in range(3):
                plan.apply(W, Y)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); plan.apply(W, Y); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = statistics.median(ts)
            gbs = info["algorithmic_bytes"] / ms / 1e6
            row = {"n": n, "N": N, "M": M, "p": p, "L": L, "ms": ms, "gdof_s": N * L / ms / 1e6, "alg_gbs": gbs,
                   "frac_hbm": gbs / peak, "kernel": plan.kernel_name}
            rows.append(row)
            print("N=%8d L=%5d M=%2d  %8.3f ms  %6.1f GDoF/s  %6.0f GB/s  frac %.3f  %s"
                  % (N, L, M, ms, row["gdof_s"], gbs, row["frac_hbm"], plan.kernel_name), flush=True)
            del A, W, Y, plan, rowptr, colidx, vals
            torch.cuda.empty_cache()
    with open(args.out, "w") as f:
        json.dump({"peak_gbs": peak, "points": rows}, f, indent=1)


if __name__ == "__main__":
    main()
