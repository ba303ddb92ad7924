#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
for halo in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --cpu-cols 3 --halo $halo > gpurun_out/r2e_bench_${N}gpu_$halo.json 2> gpurun_out/r2e_bench_${N}gpu_$halo.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2e_bench_${N}gpu_$halo.json').read().strip().splitlines()[-1])
    print('$halo', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_relerr'), d['gpu_launches'])
except Exception as e:
    print('$halo FAILED', e); print(open('gpurun_out/r2e_bench_${N}gpu_$halo.err').read()[-1500:])
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/pcg_bench.py --config c4 > gpurun_out/r2e_pcg_c4_${N}gpu.json 2> gpurun_out/r2e_pcg_c4_${N}gpu.err
cat gpurun_out/r2e_pcg_c4_${N}gpu.json | cut -c1-420; tail -2 gpurun_out/r2e_pcg_c4_${N}gpu.err | cut -c1-400
