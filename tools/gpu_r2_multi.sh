#!/bin/bash
# N GPUs of one box: the bench line (row sharding, peer halos, with the row-sharded c4 PCG block), the NCCL-halo
# variant, and tools/pcg_bench.py on config 4
mkdir -p gpurun_out
N=${1:-8}
O=gpurun_out
run_bench() { # extra tag
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --cpu-cols 3 $1 > $O/r2_bench_c3_$2.json 2> $O/r2_bench_c3_$2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_c3_$2.json').read().strip().splitlines()[-1])
    print('$2', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d.get('parity_relerr'), d['gpu_launches'], d.get('e2e',{}).get('value'))
    if d.get('pcg'): print('  pcg', d['pcg']['numit'], round(d['pcg']['seconds'],4), round(d['pcg']['ms_per_pass'],3), d['pcg']['max_abs_error_vs_manufactured'])
except Exception as e:
    print('$2 FAILED', e); print(open('gpurun_out/r2_bench_c3_$2.err').read()[-1500:])
PY
}
run_bench "" ${N}gpu_rows
run_bench "--halo nccl --no-e2e --no-pcg" ${N}gpu_rows_nccl_halo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/pcg_bench.py --config c4 > $O/r2_pcg_c4_${N}gpu_rows.json 2> $O/r2_pcg_c4_${N}gpu_rows.err
cut -c1-420 $O/r2_pcg_c4_${N}gpu_rows.json
