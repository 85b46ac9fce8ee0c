#!/bin/bash
for cfg in "TBN_GROUP=0 TBN_N_HI=0" "TBN_GROUP=0 TBN_N_HI=3" "TBN_GROUP=0 TBN_N_HI=2"; do
  for chunk in 32 48; do
  env $cfg timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --chunk $chunk 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg chunk $chunk', 'ms/step %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_step'], 'launches/step', d['gpu_launches'] // d['steps'])"
  done
done
