#!/usr/bin/env python
"""Per-instruction stall reasons from an ncu report (development tool).
usage: prof_lines.py <report.ncu-rep> lo:hi [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
lo, hi = [int(v) for v in sys.argv[2].split(":")]
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = rows[2:]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for i in range(lo, min(hi, len(data))):
    r = data[i]
    if int(r[isamp]) < mins:
        continue
    top = sorted(((int(r[j] or 0), n) for j, n in stall), reverse=True)[:3]
    print(i, r[isamp].rjust(6), r[iexec].rjust(9), r[isrc][:60].ljust(60), " ".join("%s=%d" % (n, v) for v, n in top if v))
