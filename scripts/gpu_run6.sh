#!/bin/bash
mkdir -p gpurun_out
for cfg in "TBN_GROUP=1 TBN_TC_PDL=0" "TBN_GROUP=1 TBN_TC_PDL=1" "TBN_GROUP=0 TBN_TC_PDL=0" "TBN_GROUP=0 TBN_TC_PDL=1"; do
  env $cfg timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'launches/step', d['gpu_launches'] // d['steps'], 'mhz', d['clocks']['sm_mhz'])"
done
