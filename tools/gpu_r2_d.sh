#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1; tail -4 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --config c3u --no-pcg --no-e2e --steps 5 --cpu-cols 3 > gpurun_out/r2d_bench_c3u.json 2> gpurun_out/r2d_bench_c3u.err; cat gpurun_out/r2d_bench_c3u.json | head -c 1800; echo
timeout 900 python bench.py > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err; cat gpurun_out/r2d_bench_c3.json | head -c 3500; echo; tail -3 gpurun_out/r2d_bench_c3.err
