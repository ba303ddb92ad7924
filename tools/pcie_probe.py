#!/usr/bin/env python
"""Pinned-memory PCIe bandwidth on this box: H2D alone, D2H alone, both at once (development tool;
the e2e leg of bench.py is bound by the last figure)."""
import torch, time
n = 1 << 30  # 1 GiB chunks
h_in = torch.empty(4 * n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(4 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(4 * n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(4 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunk):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for o in range(0, 4 * n, chunk):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    return 4 * n / (time.perf_counter() - t) / 1e9
for chunk in (4 * n, n, n // 8):
    for _ in range(2):
        a, b, c = run(True, False, chunk), run(False, True, chunk), run(True, True, chunk)
    print("chunk %5d MiB: H2D %.1f GB/s, D2H %.1f GB/s, both %.1f GB/s each direction" % (chunk >> 20, a, b, c), flush=True)
