#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/exp/bin/tmem_drain > gpurun_out/tmem_drain.txt 2>&1; cat gpurun_out/tmem_drain.txt
timeout 600 python -m pytest tests/test_parity_bf16_gpu.py tests/test_model_gpu.py -q -m gpu --timeout 180 -s 2>&1 | tee gpurun_out/tests2.log | grep -E "cfg5|passed|failed|rror" | tail -20
