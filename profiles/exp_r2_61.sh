# two-contact islands: body velocities forwarded between the two contacts in registers instead of through the lane's shared-memory slots
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
OLD=gym_puzzles_b200/csrc/build/var/libmrp_old.so
for i in 1 2 3; do echo "== old"; MRP_LIB_PATH=$OLD python profiles/quickbench.py; echo "== new"; python profiles/quickbench.py; done
echo "== 262144 old, new"; MRP_LIB_PATH=$OLD QB_ENVS=262144 python profiles/quickbench.py; QB_ENVS=262144 python profiles/quickbench.py
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for L in old new; do
P=""; [ $L = old ] && P="MRP_LIB_PATH=$OLD"
env $P QB_ENVS=1048576 MRP_SPARES=0 ncu --metrics $M --clock-control none -k 'regex:^k_solve_vel$' -s 62 -c 1 --csv --log-file gpurun_out/r2_exp61_$L.csv python profiles/profile_step.py > gpurun_out/r2_exp61_$L.log 2>&1
done
