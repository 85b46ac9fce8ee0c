#!/bin/bash
# GPU-box check: kernel tests, model tests (per-test timeouts so that nothing can hang the call), short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$1" != "bench" ]; then
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 60 -x 2>&1 | tee gpurun_out/kernels.log | tail -40
timeout 420 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 120 2>&1 | tee gpurun_out/model.log | tail -60
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
