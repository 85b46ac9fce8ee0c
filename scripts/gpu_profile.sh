#!/bin/bash
# ncu launch list of one forward (run plain first, then under ncu, per B200_PROFILING.md).
mkdir -p gpurun_out
N=${1:-356}
python scripts/profile_step.py > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:avgpool_feat|consensus|conv_tc|groupnorm|im2col|linear_kernel|maxpool3s2|mha_core|pool3s1|conv_simt|nchw_to|attn_weighted|softmax_rows" -s $N -c $N --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py > gpurun_out/profile_ncu.log 2>&1
cat gpurun_out/profile_plain.log | tail -2
