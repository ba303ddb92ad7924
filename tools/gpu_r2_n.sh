#!/bin/bash
# spike correction folded into phase 3, CPT=2 separator kernel without local memory; full GPU suite
O=gpurun_out
mkdir -p $O
for g in "500 500 5000" "512 512 1771" "1000 1000 988"; do set -- $g
  timeout 300 python tools/precond_bench.py --nx $1 --ny $2 --L $3 2>&1 | tail -1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file $O/r2n_precond_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 2 > $O/r2n_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2n_precond_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[ki].split('(')[0]; tot[n][0]+=1; tot[n][1]+=float(r[vi].replace(',',''))/1e6
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:8]: print(n[:70],c,round(t/c,3))
PY
timeout 900 python tools/pcg_bench.py --config c4 > $O/r2n_pcg_c4.json 2> $O/r2n_pcg_c4.err; cut -c1-420 $O/r2n_pcg_c4.json; tail -2 $O/r2n_pcg_c4.err | cut -c1-300
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2n_pytest.log 2>&1; tail -5 $O/r2n_pytest.log
