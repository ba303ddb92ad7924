#!/bin/bash
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-pcg > $O/r2_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[ki].split('(')[0]; tot[n][0]+=1; tot[n][1]+=float(r[vi].replace(',',''))/1e6
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:6]: print(n[:60],c,round(t,2))
PY
