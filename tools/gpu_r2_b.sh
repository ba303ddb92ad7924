#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "mean or lu or variable" > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
python tools/precond_bench.py --nx 500 --ny 500 --L 5000 > gpurun_out/r2b_precond.json 2>gpurun_out/r2b_precond.err
cat gpurun_out/r2b_precond.json
python tools/precond_bench.py --nx 512 --ny 512 --L 1771 >> gpurun_out/r2b_precond.json 2>>gpurun_out/r2b_precond.err
python tools/precond_bench.py --nx 1000 --ny 1000 --L 988 >> gpurun_out/r2b_precond.json 2>>gpurun_out/r2b_precond.err
tail -2 gpurun_out/r2b_precond.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_precond_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1 > gpurun_out/r2b_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2b_precond_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
for r in rows[1:]:
    if 'spuq' in r[ki]: print(r[ki][:60], r[vi])
PY
