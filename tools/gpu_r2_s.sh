#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python tools/pcg_bench.py --config c4 2> $O/r2s_a.err | cut -c1-330
SPUQ_PCG_NO_PREFUSE=1 timeout 900 python tools/pcg_bench.py --config c4 2> $O/r2s_b.err | cut -c1-330
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/r2s_pcg_launches.csv python tools/pcg_bench.py --config c4 > $O/r2s_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2s_pcg_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.defaultdict(lambda:[0,0.0,0.0])
for r in rows[1:]:
    n=r[ki].split('(')[0]; t=float(r[vi].replace(',',''))/1e6; tot[n][0]+=1; tot[n][1]+=t; tot[n][2]=max(tot[n][2],t)
for n,(c,t,m) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:8]: print(n[:70],c,'avg',round(t/c,3),'max',round(m,3))
PY
