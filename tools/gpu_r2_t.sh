#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_schur -s 1 -c 1 -o $O/r2t_schur_full -f python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1 > $O/r2t_ncu_full.log 2>&1
ls -la $O/r2t_schur_full.ncu-rep
