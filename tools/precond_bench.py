#!/usr/bin/env python
"""Time the exact mean preconditioner on a slab:  python tools/precond_bench.py --nx 500 --ny 500 --L 5000"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=500)
    ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--L", type=int, default=5000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--lib", default=None, help="another build of libspuq_sg.so (A/B runs of compile-time switches)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    from spuq_b200 import _abi
    if a.lib:
        _abi.use_library(os.path.abspath(a.lib), "cuda:0")
    from spuq_b200.multi_operator import MeanSolveSlabOperator
    dev = _abi.device()
    N = a.nx * a.ny
    op = MeanSolveSlabOperator(None, a.L, stencil=(a.nx, a.ny, 4.0, -1.0, -1.0))
    R = torch.rand((a.L, N), dtype=torch.float64, device=dev)
    S = torch.empty_like(R)
    op.apply_slab(R, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        op.apply_slab(R, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    # residual check on one column: A0 s = r
    s = S[0].view(a.ny, a.nx); r = R[0].view(a.ny, a.nx)
    As = 4.0 * s
    As[:, 1:] -= s[:, :-1]; As[:, :-1] -= s[:, 1:]; As[1:, :] -= s[:-1, :]; As[:-1, :] -= s[1:, :]
    res = float((As - r).abs().max() / r.abs().max())
    print(json.dumps({"lib": os.path.basename(a.lib) if a.lib else "default", "nx": a.nx, "ny": a.ny, "L": a.L, "ms": ms, "GB_per_s_24NL": 24.0 * N * a.L / ms / 1e6,
                      "residual": res}))


if __name__ == "__main__":
    main()
