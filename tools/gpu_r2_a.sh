#!/bin/bash
# round-2 GPU check A: GPU test suite + config-4 PCG with the strip preconditioner (and the plain transform for comparison)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python tools/pcg_bench.py --config c4 > gpurun_out/r2a_pcg_c4.json 2> gpurun_out/r2a_pcg_c4.err
cat gpurun_out/r2a_pcg_c4.json | head -c 1500; echo
SPUQ_PRECOND_PLAIN=1 timeout 600 python tools/pcg_bench.py --config c4 > gpurun_out/r2a_pcg_c4_plain.json 2> gpurun_out/r2a_pcg_c4_plain.err
cat gpurun_out/r2a_pcg_c4_plain.json | head -c 600; echo
