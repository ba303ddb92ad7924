#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "general_sparse or unstructured or apply_paths" > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
for c in c3n c3u; do
  timeout 900 python bench.py --config $c --no-pcg --no-e2e --steps 10 --cpu-cols 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_$c.json').read().strip().splitlines()[-1]); print('$c', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['parity_relerr'], d['details']['kernel'])"
done
ncu --set full --clock-control none --import-source on -k regex:k_apply_ellblk -s 3 -c 1 -f -o gpurun_out/r2_c3n_ellblk python bench.py --config c3n --steps 1 --warmup 3 --no-e2e --no-cpu --no-pcg > gpurun_out/r2_ncu_full_c3n.log 2>&1
python tools/ncu_to_json.py gpurun_out/r2_c3n_ellblk.ncu-rep gpurun_out/r2_c3n_ellblk_ncu_full.json gpurun_out/r2_traffic.json "c3n|k_apply_ellblk<5>" "profiles/r2_c3n_ellblk_ncu_full.json (ncu --set full, one launch of bench.py --config c3n)" | grep -E "duration|inst_executed|issue_active|fp64"
