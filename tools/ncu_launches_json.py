#!/usr/bin/env python
"""Key metrics of EVERY kernel launch of an ncu report -> JSON (for profiles/).
usage: ncu_launches_json.py <report.ncu-rep> <out.json> "<command that produced the report>" """
import csv, io, json, subprocess, sys
rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct"]
launches = []
for vals in rows[2:]:
    d = {}
    for n, u, v in zip(names, units, vals):
        if n in want or (n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")):
            d[n] = {"value": v, "unit": u}
    launches.append(d)
json.dump({"command": cmd, "launches": launches}, open(out, "w"), indent=1)
print(len(launches), "launches ->", out)
