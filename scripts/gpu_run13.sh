#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "stem" 2>&1 | tail -3
python scripts/trace_stem.py 2>&1 | tee gpurun_out/trace_stem2.txt | grep -E "==|staged|row1 built|next tile"
python scripts/bench_stem.py 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['config'].get('launch_plan'), d['config'].get('launch_plan_ms'))"
