#!/usr/bin/env python
"""Summarise an ncu report (development tool): key metrics + stall samples by SASS region.
usage: prof_summary.py <report.ncu-rep> [region_size] [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum',
          'dram__bytes_write.sum', 'lts__t_sectors_srcunit_tex.sum', 'launch__registers_per_thread',
          'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum.per_second']:
    print(k, d.get(k))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "n instr", len(data))
for b in range(0, len(data), blk):
    ss = sum(int(r[isamp]) for r in data[b:b + blk])
    ex = max(int(r[iexec]) for r in data[b:b + blk])
    ops = {}
    for r in data[b:b + blk]:
        t = r[isrc].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    if ss > mins:
        print(b, ss, "%.1f%%" % (100.0 * ss / tot), ex, top)
if len(sys.argv) > 4:
    lo, hi = [int(v) for v in sys.argv[4].split(":")]
    for i in range(lo, hi):
        r = data[i]
        print(i, r[isamp].rjust(6), r[iexec].rjust(9), r[isrc][:110])
