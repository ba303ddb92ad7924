#!/bin/bash
# launch list of the c4 solve (per-kernel times inside the PCG loop)
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/r2l_pcg_launches.csv python tools/pcg_bench.py --config c4 > $O/r2l_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2l_pcg_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[ki].split('(')[0]; tot[n][0]+=1; tot[n][1]+=float(r[vi].replace(',',''))/1e6
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:14]: print(n[:70],c,round(t/c,3), round(t,1))
PY
