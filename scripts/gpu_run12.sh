#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 240 -x 2>&1 | tail -4
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
python scripts/profile_step.py --cfg 3 --group 0 > gpurun_out/prof_plain_cfg3.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:"attn_weighted|consensus|avgpool_feat" --csv --log-file gpurun_out/ncu_mem_cfg3.csv python scripts/profile_step.py --cfg 3 --group 0 --reps 2 > /dev/null 2>&1
python scripts/profile_step.py --cfg 3 --group 0 --clips 256 --reps 1 > /dev/null 2>&1 && \
ncu --metrics $M --clock-control none -k regex:"attn_weighted|consensus|avgpool_feat" --csv --log-file gpurun_out/ncu_mem_cfg3_b256.csv python scripts/profile_step.py --cfg 3 --group 0 --clips 256 --reps 2 > /dev/null 2>&1
wc -l gpurun_out/ncu_mem_cfg3.csv gpurun_out/ncu_mem_cfg3_b256.csv
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2b.json 2>gpurun_out/bench_r2b.err; tail -1 gpurun_out/bench_r2b.json | cut -c1-200; grep -o '"roofline.*' gpurun_out/bench_r2b.json | cut -c1-700
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-900
