#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "group or conv_tc" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 240 -x -s -k "grouped or golden" 2>&1 | grep -E "launches|passed|failed|rror" | tail -12
timeout 300 python scripts/layer_times.py > gpurun_out/layer_times_group.txt 2>&1; tail -16 gpurun_out/layer_times_group.txt
TBN_GROUP=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nogroup.json 2> gpurun_out/bench_nogroup.err; tail -2 gpurun_out/bench_nogroup.err; cut -c1-400 gpurun_out/bench_nogroup.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_group.json 2> gpurun_out/bench_group.err; tail -2 gpurun_out/bench_group.err; cut -c1-400 gpurun_out/bench_group.json; grep -o '"roofline".*' gpurun_out/bench_group.json | cut -c1-300
