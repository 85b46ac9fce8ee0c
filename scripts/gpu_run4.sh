#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x 2>&1 | tail -5
timeout 300 python scripts/layer_times.py --tune tc_epi=1 > gpurun_out/layer_times_epi1.txt 2>&1; tail -14 gpurun_out/layer_times_epi1.txt
timeout 300 python scripts/layer_times.py --tune tc_epi=0 --quiet 1 > gpurun_out/layer_times_epi0.txt 2>&1; tail -3 gpurun_out/layer_times_epi0.txt
timeout 300 python scripts/layer_times.py --tune tc_epi=1 tc_cta2=1 --quiet 0 > gpurun_out/layer_times_epi1_cta2_1.txt 2>&1; tail -2 gpurun_out/layer_times_epi1_cta2_1.txt
timeout 300 python scripts/layer_times.py --tune tc_epi=1 tc_cta2=2 --quiet 0 > gpurun_out/layer_times_epi1_cta2_2.txt 2>&1; tail -2 gpurun_out/layer_times_epi1_cta2_2.txt
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_parity_bf16_gpu.py -q -m gpu --timeout 180 -x 2>&1 | tail -5
