#!/bin/bash
# producer back-off of k_apply_tmem: sustained bench time against the sleep length
O=gpurun_out
mkdir -p $O
for k in 0 16 32 48 64; do
  SPUQ_KDEBUG=$k timeout 600 python bench.py --no-e2e --no-cpu --no-pcg --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('KDEBUG=$k', round(d['value'],2), round(d['ms_per_step'],4), d['clocks'])"
done
