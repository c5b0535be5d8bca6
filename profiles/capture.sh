#!/bin/bash
# ncu passes of one steady-state step (run under gpurun, 1 GPU):  bash profiles/capture.sh <tag>
# 1) launch list (gpu__time_duration per launch, last 3 steps)  2) --set full of every kernel of the last step
set -e
TAG=${1:-rX}
CMD="python profiles/profile_step.py"
KERNELS='regex:^k_(broad|narrow|pre|solve_vel|solve_pos|post|post_events|reset_list)$'
$CMD > gpurun_out/plain_$TAG.log 2>&1
# 1 reset + 64 steps x 8 matching kernels: skip 61 steps
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 488 -c 24 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "$KERNELS" -s 504 -c 8 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log
