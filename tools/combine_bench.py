#!/usr/bin/env python
"""Streaming rate of spuq_sg_combine (expansion evaluated at S parameter samples) on one GPU:
  python tools/combine_bench.py [--n 1000000 --L 988]
Prints ms and GB/s (8 n L bytes per pass of up to 8 samples) for S = 1, 4, 8, 16."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--L", type=int, default=988)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    from spuq_b200 import _abi
    lib, dev = _abi.lib(), _abi.device()
    n, L = args.n, args.L
    W = torch.rand((L, n), dtype=torch.float64, device=dev)
    res = []
    for S in (1, 4, 8, 16):
        coef = torch.rand((S, L), dtype=torch.float64, device=dev)
        out = torch.empty((S, n), dtype=torch.float64, device=dev)
        fn = lambda: _abi.check(lib.spuq_sg_combine(n, L, S, _abi.ptr(W), n, _abi.ptr(coef), _abi.ptr(out), n,  # noqa: E731
                                                    _abi.stream_ptr()), "combine")
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        passes = (S + 7) // 8
        ref = coef @ W
        err = float((out - ref).abs().max() / ref.abs().max())
        res.append({"S": S, "ms": ms, "GB/s": 8.0 * n * L * passes / ms / 1e6, "rel_err_vs_torch": err})
        print(res[-1], flush=True)
    print(json.dumps({"n": n, "L": L, "results": res}))


if __name__ == "__main__":
    main()
