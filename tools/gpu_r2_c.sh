#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "mean_precond or lu or variable" > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
python tools/precond_bench.py --nx 500 --ny 500 --L 5000 > gpurun_out/r2c_precond.json 2>gpurun_out/r2c_precond.err
cat gpurun_out/r2c_precond.json
ncu --set full --clock-control none --import-source on -k regex:k_strip -c 2 -o gpurun_out/r2c_strip -f python tools/precond_bench.py --nx 500 --ny 500 --L 600 --reps 1 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
