#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "head_chain or groupnorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_bf16_gpu.py tests/test_evaluation.py -q -m gpu --timeout 240 -x 2>&1 | tail -5
for cfg in "TBN_GROUP=0" "TBN_GROUP=1"; do
env $cfg timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/bh.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'launches/step', d['gpu_launches'] // d['steps'], d['config'].get('launch_plan'))"
done; tail -3 gpurun_out/bh.err
