#!/bin/bash
for chunk in 32 48 64 96; do
  TBN_GROUP=0 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --chunk $chunk 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk $chunk', 'ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'launches/step', d['gpu_launches'] // d['steps'])"
done
