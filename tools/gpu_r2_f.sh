#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/pcg_bench.py --config c4 > gpurun_out/r2f_pcg_c4_${N}gpu.json 2> gpurun_out/r2f_pcg_c4_${N}gpu.err
cat gpurun_out/r2f_pcg_c4_${N}gpu.json | cut -c1-900; tail -3 gpurun_out/r2f_pcg_c4_${N}gpu.err | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 --cpu-cols 3 > gpurun_out/r2f_bench_${N}gpu.json 2> gpurun_out/r2f_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_${N}gpu.json').read().strip().splitlines()[-1])
print($N, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_relerr'), d['gpu_launches'], d['e2e']['value'])
PY
