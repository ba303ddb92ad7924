#!/bin/bash
# 8-GPU box: row-sharded bench (overlap / no overlap), M-sharded bench, row-sharded PCG of config 4
mkdir -p gpurun_out
run_bench() { # N shard extra tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $1 --steps 20 --warmup 3 --cpu-cols 3 --shard $2 $3 > gpurun_out/r2g_bench_$4.json 2> gpurun_out/r2g_bench_$4.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2g_bench_$4.json').read().strip().splitlines()[-1])
    print('$4', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d.get('parity_relerr'), d['gpu_launches'], d.get('e2e',{}).get('value'))
except Exception as e:
    print('$4 FAILED', e)
PY
}
run_bench 8 rows "" 8gpu_rows
run_bench 8 rows "--no-overlap --no-e2e" 8gpu_rows_noverlap
run_bench 4 rows "--no-e2e" 4gpu_rows
run_bench 2 m "--no-e2e --no-cpu" 2gpu_m
run_bench 4 m "--no-e2e --no-cpu" 4gpu_m
run_bench 8 m "--no-e2e --no-cpu" 8gpu_m
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/pcg_bench.py --config c4 > gpurun_out/r2g_pcg_c4_${N}gpu.json 2> gpurun_out/r2g_pcg_c4_${N}gpu.err
cat gpurun_out/r2g_pcg_c4_${N}gpu.json | cut -c1-420
done
