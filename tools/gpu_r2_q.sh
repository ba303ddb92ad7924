#!/bin/bash
# N GPUs: multi-GPU tests, bench line with the row-sharded PCG block, sharded c4 solve
mkdir -p gpurun_out
N=${1:-2}
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/r2q_pytest.log 2>&1; tail -3 $O/r2q_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2q_bench_${N}gpu.json 2> $O/r2q_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2q_bench_${N}gpu.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d.get('parity_relerr'), d['gpu_launches'], d['e2e']['value'])
    print(json.dumps(d.get('pcg'))[:600])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2q_bench_${N}gpu.err').read()[-2500:])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/pcg_bench.py --config c4 > $O/r2q_pcg_c4_${N}gpu.json 2> $O/r2q_pcg_c4_${N}gpu.err
cut -c1-420 $O/r2q_pcg_c4_${N}gpu.json; tail -2 $O/r2q_pcg_c4_${N}gpu.err | cut -c1-400
