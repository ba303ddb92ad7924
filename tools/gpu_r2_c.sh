#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "mean_precond or lu or variable or pcg_sg_mean" > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
python tools/precond_bench.py --nx 500 --ny 500 --L 5000 > gpurun_out/r2c_precond.json 2>gpurun_out/r2c_precond.err
SPUQ_STRIP_GENERIC=1 python tools/precond_bench.py --nx 500 --ny 500 --L 5000 >> gpurun_out/r2c_precond.json 2>>gpurun_out/r2c_precond.err
python tools/precond_bench.py --nx 512 --ny 512 --L 1771 >> gpurun_out/r2c_precond.json 2>>gpurun_out/r2c_precond.err
python tools/precond_bench.py --nx 1000 --ny 1000 --L 988 >> gpurun_out/r2c_precond.json 2>>gpurun_out/r2c_precond.err
cat gpurun_out/r2c_precond.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c_precond_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1 > gpurun_out/r2c_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c_precond_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
for r in rows[1:16]:
    if 'spuq' in r[ki]: print(r[ki][:70], r[vi])
PY
timeout 600 python bench.py --config c3n --no-pcg --no-e2e --steps 5 --cpu-cols 3 > gpurun_out/r2c_bench_c3n.json 2> gpurun_out/r2c_bench_c3n.err; python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_c3n.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_relerr'], d['details'])"
