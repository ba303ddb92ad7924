#!/bin/bash
# Round-2 evidence on ONE B200: GPU tests, the bench line, launch lists and ncu captures (see profiles/README.md).
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2_pytest_gpu.log 2>&1; tail -4 $O/r2_pytest_gpu.log
timeout 900 python bench.py > $O/r2_bench_c3_1gpu.json 2> $O/r2_bench_c3_1gpu.err; head -c 600 $O/r2_bench_c3_1gpu.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_c3_reference_arm.json 2> $O/r2_bench_c3_reference_arm.err; head -c 400 $O/r2_bench_c3_reference_arm.json; echo
for c in c3u c3n c2 c4; do
  timeout 900 python bench.py --config $c --no-pcg --no-e2e --steps 10 --cpu-cols 3 > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err
  python -c "
import json; d=json.loads(open('$O/r2_bench_$c.json').read().strip().splitlines()[-1]); print('$c', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['parity_relerr'], d['details']['kernel'])"
done
# launch list of the bench command (kernel share of the step) and of one config-4 solve
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-pcg > $O/r2_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2_pcg_c4_launches.csv python tools/pcg_bench.py --config c4 > $O/r2_ncu_pcg.log 2>&1
python - <<'PY'
import csv, collections
for f, out in (('gpurun_out/r2_pcg_c4_launches.csv', 'gpurun_out/r2_pcg_c4_launch_summary.txt'),):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        n = r[ki].split('(')[0]
        tot[n][0] += 1; tot[n][1] += float(r[vi].replace(',', '')) / 1e6
    total = sum(v[1] for v in tot.values())
    with open(out, 'w') as fh:
        fh.write('kernel, launches, total ms, share (ncu launch list of tools/pcg_bench.py --config c4; per-launch times are cold and serialised)\n')
        for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            fh.write('%s, %d, %.2f, %.1f%%\n' % (n, c, t, 100 * t / total))
    print(open(out).read()[:1500])
PY
# full captures: the apply kernel on c3, the strip kernels, the general sparse kernel
ncu --set full --clock-control none --import-source on -k regex:k_apply_tmem -s 3 -c 1 -f -o $O/r2_c3_apply_tmem python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-pcg > $O/r2_ncu_full_c3.log 2>&1
python tools/ncu_to_json.py $O/r2_c3_apply_tmem.ncu-rep $O/r2_c3_apply_tmem_ncu_full.json $O/r2_traffic.json "c3|k_apply_tmem<K=2,G=2,5pt>" "profiles/r2_c3_apply_tmem_ncu_full.json (ncu --set full --clock-control none, one launch of bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-pcg)" > /dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_strip|k_schur" -s 3 -c 3 -f -o $O/r2_strip python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1 > $O/r2_ncu_full_strip.log 2>&1
python tools/ncu_lines.py $O/r2_strip.ncu-rep k_strip 0 22 > $O/r2_precond_strip_lines.txt; python tools/ncu_lines.py $O/r2_strip.ncu-rep k_schur 0 12 >> $O/r2_precond_strip_lines.txt; python tools/ncu_lines.py $O/r2_strip.ncu-rep k_strip 1 22 >> $O/r2_precond_strip_lines.txt
python tools/ncu_launches_json.py $O/r2_strip.ncu-rep $O/r2_precond_strip_ncu_full.json "ncu --set full --clock-control none --import-source on -k regex:k_strip|k_schur -s 3 -c 3 python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file $O/r2_precond_strip_launches.csv python tools/precond_bench.py --nx 500 --ny 500 --L 5000 --reps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply_ellblk -s 3 -c 1 -f -o $O/r2_c3n_ellblk python bench.py --config c3n --steps 1 --warmup 3 --no-e2e --no-cpu --no-pcg > $O/r2_ncu_full_c3n.log 2>&1
python tools/ncu_to_json.py $O/r2_c3n_ellblk.ncu-rep $O/r2_c3n_ellblk_ncu_full.json $O/r2_traffic.json "c3n|k_apply_ellblk<5>" "profiles/r2_c3n_ellblk_ncu_full.json (ncu --set full, one launch of bench.py --config c3n)" > /dev/null
for s in "500 500 5000" "512 512 1771" "1000 1000 988"; do set -- $s; python tools/precond_bench.py --nx $1 --ny $2 --L $3; done > $O/r2_precond_bench.json 2>/dev/null; cat $O/r2_precond_bench.json
ls -la $O | tail -30
