#!/bin/bash
# GPU-box check: all GPU tests (per-test timeouts so that nothing can hang the call), then a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$1" != "bench" ]; then
timeout 900 python -m pytest tests -q -m gpu --timeout 180 -x -s 2>&1 | tee gpurun_out/tests.log | grep -E "parity|top-1|passed|failed|error|Error" | tail -80
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
