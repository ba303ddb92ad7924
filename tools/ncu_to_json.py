#!/usr/bin/env python
"""Key metrics of one kernel launch of an ncu report -> JSON (for profiles/), plus the DRAM bytes
entry that bench.py reads for roofline.traffic.
usage: ncu_to_json.py <report.ncu-rep> <out.json> [traffic.json traffic_key source_note]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum.per_second"]
d = {}
for n, u, v in zip(names, units, vals):
    if n in want or n.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in n:
        d[n] = {"value": v, "unit": u}
json.dump(d, open(out, "w"), indent=1)
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
if len(sys.argv) > 5:
    rd = float(d["dram__bytes_read.sum"]["value"].replace(",", "")) * scale[d["dram__bytes_read.sum"]["unit"]]
    wr = float(d["dram__bytes_write.sum"]["value"].replace(",", "")) * scale[d["dram__bytes_write.sum"]["unit"]]
    tj = sys.argv[3]
    try:
        t = json.load(open(tj))
    except Exception:
        t = {}
    t[sys.argv[4]] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "source": sys.argv[5]}
    json.dump(t, open(tj, "w"), indent=1)
print(json.dumps({k: v["value"] for k, v in d.items() if "pcsamp" not in k}, indent=1))
