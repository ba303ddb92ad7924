#!/bin/bash
# 8-GPU box: row-sharded bench (peer / nccl halo exchange) and row-sharded PCG of config 4
mkdir -p gpurun_out
run_bench() { # N extra tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $1 --steps 20 --warmup 3 --cpu-cols 3 $2 > gpurun_out/r2g_bench_$3.json 2> gpurun_out/r2g_bench_$3.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2g_bench_$3.json').read().strip().splitlines()[-1])
    print('$3', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d.get('parity_relerr'), d['gpu_launches'], d.get('e2e',{}).get('value'))
except Exception as e:
    print('$3 FAILED', e); print(open('gpurun_out/r2g_bench_$3.err').read()[-800:])
PY
}
run_bench 8 "" 8gpu_rows_peer
run_bench 8 "--halo nccl --no-e2e" 8gpu_rows_nccl
run_bench 4 "--no-e2e" 4gpu_rows_peer
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/pcg_bench.py --config c4 > gpurun_out/r2g_pcg_c4_${N}gpu.json 2> gpurun_out/r2g_pcg_c4_${N}gpu.err
cat gpurun_out/r2g_pcg_c4_${N}gpu.json | cut -c1-420
done
