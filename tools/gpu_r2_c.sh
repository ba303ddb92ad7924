#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "mean_precond or lu or variable or pcg_sg_mean or partitioned" > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
for s in "500 500 5000" "512 512 1771" "1000 1000 988"; do set -- $s; python tools/precond_bench.py --nx $1 --ny $2 --L $3; done > gpurun_out/r2_precond_bench.json 2>/dev/null; cat gpurun_out/r2_precond_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c_precond_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1 > gpurun_out/r2c_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c_precond_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
for r in rows[1:16]:
    if 'spuq' in r[ki]: print(r[ki][:70], r[vi])
PY
ncu --set full --clock-control none --import-source on -k regex:k_strip -c 4 -f -o gpurun_out/r2_strip python tools/precond_bench.py --nx 500 --ny 500 --L 600 --reps 1 > gpurun_out/r2_ncu_full_strip.log 2>&1
