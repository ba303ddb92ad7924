#!/bin/bash
# last sanity pass on the final tree: smoke(), the GPU suite, the default bench line
O=gpurun_out
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > $O/r2w_bench.json 2> $O/r2w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['parity_relerr'], d['e2e']['value'], d['pcg']['seconds'], d['pcg']['launches_per_pass'], d['clocks'])
PY
