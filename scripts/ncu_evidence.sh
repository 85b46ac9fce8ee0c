#!/bin/bash
# ncu evidence of round 2 (one GPU): per-launch DRAM bytes + durations of every kernel of one cfg-2 step (both launch
# plans), the memory-bound kernels at cfg 2 / cfg 3 / a large batch, and a --set full capture of the grouped conv launches.
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
for g in 1 0; do
  python scripts/profile_step.py --group $g > gpurun_out/prof_plain_g$g.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/prof_plain_g$g.log; exit 1; }
  N=$(grep launches_per_forward gpurun_out/prof_plain_g$g.log | awk '{print $2}')
  ncu --metrics $M --clock-control none -s $N -c $N --csv --log-file gpurun_out/ncu_step_g$g.csv python scripts/profile_step.py --group $g > gpurun_out/prof_ncu_g$g.log 2>&1
  echo "group=$g launches=$N rows=$(wc -l < gpurun_out/ncu_step_g$g.csv)"
done
# memory-bound kernels: fixed-attention config (attn_weighted_sum) and a large batch (256 clips = 768 frames)
python scripts/profile_step.py --cfg 3 --group 0 > gpurun_out/prof_plain_cfg3.log 2>&1
N=$(grep launches_per_forward gpurun_out/prof_plain_cfg3.log | awk '{print $2}')
ncu --metrics $M --clock-control none -s $N -c $N -k regex:"pool|consensus|attn_weighted|mha_core|groupnorm|split_bf16|splitk_reduce" --csv --log-file gpurun_out/ncu_mem_cfg3.csv python scripts/profile_step.py --cfg 3 --group 0 > /dev/null 2>&1
python scripts/profile_step.py --clips 256 --group 0 > gpurun_out/prof_plain_b256.log 2>&1
N=$(grep launches_per_forward gpurun_out/prof_plain_b256.log | awk '{print $2}')
ncu --metrics $M --clock-control none -k regex:"pool|consensus|attn_weighted|mha_core|groupnorm|split_bf16|splitk_reduce" -s 0 -c 400 --csv --log-file gpurun_out/ncu_mem_b256.csv python scripts/profile_step.py --clips 256 --group 0 --reps 2 > /dev/null 2>&1
echo "mem rows: $(wc -l < gpurun_out/ncu_mem_cfg3.csv) $(wc -l < gpurun_out/ncu_mem_b256.csv)"
# --set full on the grouped conv launches of inception 4a-4e (second forward)
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 66 -c 12 -o gpurun_out/r2_conv_tc_grouped -f python scripts/profile_step.py --group 1 > gpurun_out/prof_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
