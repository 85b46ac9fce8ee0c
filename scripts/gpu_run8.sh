#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "stem" 2>&1 | tail -3
for pf in 0 1; do
timeout 300 python scripts/layer_times.py --tune stem_prefetch=$pf --reps 3 2>&1 | grep -E "tbn_stem_fused_tc  |everything" | tail -2
done
for pf in 0 1; do
 timeout 300 python scripts/layer_times.py --tune stem_prefetch=$pf --reps 1 2>&1 | grep -E "tbn_stem_fused_tc  " | tail -1
done
