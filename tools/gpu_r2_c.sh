#!/bin/bash
mkdir -p gpurun_out
python tools/precond_bench.py --nx 500 --ny 500 --L 5000 > gpurun_out/r2c_precond.json 2>gpurun_out/r2c_precond.err
cat gpurun_out/r2c_precond.json
ncu --set full --clock-control none --import-source on -k regex:k_strip -c 2 -o gpurun_out/r2c_strip python tools/precond_bench.py --nx 500 --ny 500 --L 600 --reps 1 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/
