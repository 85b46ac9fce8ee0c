#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "stem" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_parity_bf16_gpu.py tests/test_model_gpu.py -q -m gpu --timeout 180 -s 2>&1 | tee gpurun_out/tests2.log | grep -E "parity|top-1|passed|failed|rror" | tail -70
timeout 120 scripts/exp/bin/tmem_drain > gpurun_out/tmem_drain.txt 2>&1; cat gpurun_out/tmem_drain.txt
timeout 300 python scripts/layer_times.py > gpurun_out/layer_times_base.txt 2>&1; tail -12 gpurun_out/layer_times_base.txt
