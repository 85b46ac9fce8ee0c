#!/bin/bash
mkdir -p gpurun_out
for cfg in "TBN_GROUP=1 TBN_GROUP_POOL_LANE=2" "TBN_GROUP=1 TBN_GROUP_POOL_LANE=0" "TBN_GROUP=0" "TBN_GROUP=auto"; do
  env $cfg timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/b7.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'launches/step', d['gpu_launches'] // d['steps'], d['config'].get('launch_plan'), d['config'].get('launch_plan_ms'))"
  tail -2 gpurun_out/b7.err | grep -i error
done
for cfg in "TBN_GROUP=1" "TBN_GROUP=0"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --config 5 2>gpurun_out/b7.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg5 $cfg', 'ms/step %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'launches/step', d['gpu_launches'] // d['steps'])"
done
