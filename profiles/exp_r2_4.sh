python -m pytest tests/test_gpu_parity.py tests/test_identical_states.py -m gpu -x -q 2>&1 | tail -3
echo "== tiled layout, old flow"; QB_PHASES=1 python profiles/quickbench.py
python profiles/quickbench.py
echo "== front flow"; MRP_FRONT=1 QB_PHASES=1 python profiles/quickbench.py
MRP_FRONT=1 python profiles/quickbench.py
echo "== 262144"; QB_ENVS=262144 python profiles/quickbench.py
QB_E2E=1 python profiles/quickbench.py
ALL=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
