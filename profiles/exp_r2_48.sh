# long velocity solves parked after 16 sweeps by k_solve_vel, finished in full warps by k_solve_vel_long (MRP_LONG=0: as before)
# (the code under test was measured and NOT kept: DESIGN.md §8, "Measured dead ends"; this script is the record of the A/B)
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do echo "== MRP_LONG=0"; MRP_LONG=0 python profiles/quickbench.py; echo "== MRP_LONG=1"; python profiles/quickbench.py; done
echo "== phases off / on"; MRP_LONG=0 QB_PHASES=1 python profiles/quickbench.py; QB_PHASES=1 python profiles/quickbench.py
echo "== 262144 / 524288 off, on"
MRP_LONG=0 QB_ENVS=262144 python profiles/quickbench.py; QB_ENVS=262144 python profiles/quickbench.py
MRP_LONG=0 QB_ENVS=524288 python profiles/quickbench.py; QB_ENVS=524288 python profiles/quickbench.py
echo "== v0 / v2 / Heavy-v2 off, on"
MRP_LONG=0 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2 MultiRobotPuzzleHeavy-v2; python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2 MultiRobotPuzzleHeavy-v2
