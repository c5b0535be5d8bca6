python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo "== vel split on"; python profiles/quickbench.py; python profiles/trace_step.py | tail -1
echo "== vel split off"; MRP_VEL_SPLIT=0 python profiles/quickbench.py; MRP_VEL_SPLIT=0 python profiles/trace_step.py | tail -1
echo "== 262144 on/off"; QB_ENVS=262144 python profiles/quickbench.py; MRP_VEL_SPLIT=0 QB_ENVS=262144 python profiles/quickbench.py
ALL=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
