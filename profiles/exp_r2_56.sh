# narrow-phase queue in two groups (octagon-octagon pairs from the front, the 8 x 4 / 4 x 4 pairs from the back): SAT loops of one length per warp
# (the code under test was measured and NOT kept: DESIGN.md §8, "Measured dead ends"; this script is the record of the A/B)
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_square_variant.py -m gpu -x -q 2>&1 | tail -2
OLD=gym_puzzles_b200/csrc/build/var/libmrp_old.so
for i in 1 2 3; do echo "== old"; MRP_LIB_PATH=$OLD python profiles/quickbench.py; echo "== new"; python profiles/quickbench.py; done
echo "== v0 / v2 old, new"
MRP_LIB_PATH=$OLD python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2; python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for L in old new; do
P=""; [ $L = old ] && P="MRP_LIB_PATH=$OLD"
env $P QB_ENVS=1048576 MRP_SPARES=0 ncu --metrics $M --clock-control none -k 'regex:^k_(narrow|broad)$' -s 124 -c 2 --csv --log-file gpurun_out/r2_exp56_$L.csv python profiles/profile_step.py > gpurun_out/r2_exp56_$L.log 2>&1
done
