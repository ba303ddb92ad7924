#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
for ov in "" "--no-overlap"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --cpu-cols 3 $ov > gpurun_out/r2e_bench_2gpu$ov.json 2> gpurun_out/r2e_bench_2gpu$ov.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2e_bench_2gpu$ov.json').read().strip().splitlines()[-1])
print('$ov', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_relerr'), d['gpu_launches'])
PY
done
tail -3 gpurun_out/r2e_bench_2gpu.err
