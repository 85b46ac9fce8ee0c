#!/bin/bash
# stem kernel: tests, L2-warm and cold (ncu) launch times, step time
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "stem" 2>&1 | tail -2
python scripts/bench_stem.py 2>&1 | tail -3
python scripts/profile_step.py --group 0 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stem_fused -s 12 -c 12 --csv --log-file gpurun_out/stem_cold.csv python scripts/profile_step.py --group 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/stem_cold.csv') if l.startswith('"'))]
t={}
for r in rows: t.setdefault(r['Kernel Name'][:40],[]).append(float(r['Metric Value'].replace(',','')))
for k,v in t.items(): print('cold (ncu):', k, ' '.join('%.1f'%(x/1e3 if x>1000 else x) for x in v))
PY
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['config'].get('launch_plan'), d['config'].get('launch_plan_ms'))"
