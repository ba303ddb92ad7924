#!/bin/bash
# A/B of the strip kernel's compile-time switches on the c4 grid
O=gpurun_out
mkdir -p $O
for lib in "" "--lib tools/ab/libspuq_sg_pre0.so" "--lib tools/ab/libspuq_sg_fold1.so"; do
  timeout 300 python tools/precond_bench.py --nx 500 --ny 500 --L 5000 $lib 2>&1 | tail -1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file $O/r2o_precond_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 2 > $O/r2o_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2o_precond_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[ki].split('(')[0]; tot[n][0]+=1; tot[n][1]+=float(r[vi].replace(',',''))/1e6
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:8]: print(n[:70],c,round(t/c,3))
PY
