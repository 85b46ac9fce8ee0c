#!/bin/bash
# ncu --set full capture of a few conv_tc launches of one forward (after a plain run exited 0).
mkdir -p gpurun_out
SKIP=${1:-274}; COUNT=${2:-4}; NAME=${3:-prof_conv_tc}
python scripts/profile_step.py > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s $SKIP -c $COUNT -o gpurun_out/$NAME -f \
    python scripts/profile_step.py > gpurun_out/profile_ncu_full.log 2>&1
tail -3 gpurun_out/profile_ncu_full.log
