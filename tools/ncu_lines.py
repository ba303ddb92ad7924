#!/usr/bin/env python
"""Stall samples, shared-memory wavefronts and instruction counts of one kernel launch of an ncu report,
aggregated per CUDA SOURCE LINE (needs -lineinfo and --import-source on).
usage: ncu_lines.py <report.ncu-rep> <kernel regex> [launch index = 0] [lines = 25]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rx, "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
keys = {"s": "# Samples", "w": "L1 Wavefronts Shared", "x": "L1 Wavefronts Shared Excessive", "inst": "Instructions Executed",
        "bar": "stall_barrier", "lsb": "stall_long_sb", "ssb": "stall_short_sb", "mio": "stall_mio", "wait": "stall_wait",
        "math": "stall_math"}


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


hdr, cur, agg, name = None, None, {}, ""
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        name = r[1]
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur = (int(r[0]), r[1].strip())
        agg.setdefault(cur, {k: 0.0 for k in keys})
        continue
    if cur is None or r[2] in ("...", ""):
        continue
    d = dict(zip(hdr[2:], r[2:]))
    for k, col in keys.items():
        agg[cur][k] += num(d.get(col, 0))
tot = {k: sum(a[k] for a in agg.values()) or 1.0 for k in keys}
print("kernel:", name[:120])
print("(absolute counts are sums over the listing and may count inlined lines twice; read the percentages)")
print("samples %d, shared wavefronts %d (excess %d), instructions %d" % (tot["s"], tot["w"], tot["x"], tot["inst"]))
print("%5s %-72s %6s %6s %6s %6s | stall samples: barrier long_sb short_sb mio wait math" % ("line", "source", "smpl%", "wave%", "exc%", "inst%"))
for cur in sorted(agg, key=lambda c: -agg[c]["s"])[:top]:
    a = agg[cur]
    print("%5d %-72s %6.1f %6.1f %6.1f %6.1f | %d %d %d %d %d %d" % (
        cur[0], cur[1][:72], 100 * a["s"] / tot["s"], 100 * a["w"] / tot["w"], 100 * a["x"] / tot["x"], 100 * a["inst"] / tot["inst"],
        a["bar"], a["lsb"], a["ssb"], a["mio"], a["wait"], a["math"]))
