#!/bin/bash
# warp-local phase 2 of the strip kernel, pipelined separator kernel
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "precond or pcg or mean" > $O/r2u_pytest.log 2>&1; tail -2 $O/r2u_pytest.log
for g in "500 500 5000" "512 512 1771" "1000 1000 988"; do set -- $g
  timeout 300 python tools/precond_bench.py --nx $1 --ny $2 --L $3 2>&1 | tail -1
done
SPUQ_STRIP_PAD=1 timeout 300 python tools/precond_bench.py --nx 500 --ny 500 --L 5000 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file $O/r2u_precond_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 2 > $O/r2u_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2u_precond_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[ki].split('(')[0]; tot[n][0]+=1; tot[n][1]+=float(r[vi].replace(',',''))/1e6
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:8]: print(n[:70],c,round(t/c,3))
PY
timeout 900 python tools/pcg_bench.py --config c4 2> $O/r2u_pcg.err | cut -c1-330
