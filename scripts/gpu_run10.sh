#!/bin/bash
for cfg in "TBN_GROUP=1 TBN_GROUP_LANES=1" "TBN_GROUP=1 TBN_GROUP_LANES=2" "TBN_GROUP=0"; do
  env $cfg timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms/step %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'kernel ms %.3f' % d['roofline']['kernel_ms_per_step'], 'launches/step', d['gpu_launches'] // d['steps'])"
done
timeout 200 python -m pytest tests/test_model_gpu.py -q -m gpu -k grouped 2>&1 | tail -2
TBN_GROUP_LANES=2 timeout 200 python -m pytest tests/test_model_gpu.py -q -m gpu -k grouped 2>&1 | tail -2
