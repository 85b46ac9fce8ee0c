#!/bin/bash
# First-contact script for a GPU box: kernel tests, model tests, short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/kernels.log
cat gpurun_out/kernels.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu 2>&1 | tail -60 > gpurun_out/model.log
cat gpurun_out/model.log
