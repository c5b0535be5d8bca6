#!/bin/bash
# ncu passes of one steady-state step (run under gpurun, 1 GPU):  bash profiles/capture.sh <tag>
# 1) launch list (gpu__time_duration per launch, last 3 steps)  2) --set full of every kernel of the last step
set -e
TAG=${1:-rX}
# spare episodes off for the step capture: their refill pass launches the same kernels a second time during the first steps after
# the reset, which would shift the skip counts below; the steady-state kernels of a step are the same either way
CMD="env MRP_SPARES=0 python profiles/profile_step.py"
# k_out_rows (row fix-up of mrp_step_host with pinned buffers, a few microseconds) is left out so that the per-step skip counts
# below stay aligned with the 11 kernels of a device-resident step
KERNELS='regex:^k_(broad|narrow|pre|solve_vel|solve_pos|solve_big|post|post_events|reset_list)$'
$CMD > gpurun_out/plain_$TAG.log 2>&1
# 64 steps x 11 matching kernels (k_post and k_post_events launch twice per step: the task-free group on the side stream,
# then the envs with solver tasks): skip 61 steps for the launch list, 63 for the full capture
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 671 -c 33 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "$KERNELS" -s 693 -c 11 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log
if [ -n "$BENCH_LIST" ]; then
# 3) launch list of the bench command itself (short form: 524288 envs, 60 settle + 3 warm-up + 3 timed steps), every launch
BCMD="python bench.py --steps 3 --warmup 3 --settle 60 --envs 524288 --no-cpu-baseline --e2e-steps 2"
$BCMD > gpurun_out/plain_bench_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" --csv --log-file gpurun_out/launches_bench_$TAG.csv $BCMD > gpurun_out/ncu3_$TAG.log 2>&1
fi
