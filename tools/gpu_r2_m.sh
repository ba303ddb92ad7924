#!/bin/bash
# CPT=2 separator kernel timing + ncu full capture of the forward strip kernel and the separator kernel
O=gpurun_out
mkdir -p $O
timeout 300 python tools/precond_bench.py --nx 500 --ny 500 --L 5000 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file $O/r2m_precond_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 2 > $O/r2m_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2m_precond_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[ki].split('(')[0]; tot[n][0]+=1; tot[n][1]+=float(r[vi].replace(',',''))/1e6
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:8]: print(n[:70],c,round(t/c,3))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 2 -c 2 -o $O/r2m_strip_full -f python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1 > $O/r2m_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_schur -s 1 -c 1 -o $O/r2m_schur_full -f python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1 > $O/r2m_ncu_full2.log 2>&1
ls -la $O/*.ncu-rep
