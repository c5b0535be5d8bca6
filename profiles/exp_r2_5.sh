python -m pytest tests/test_gpu_parity.py tests/test_identical_states.py -m gpu -x -q 2>&1 | tail -3
echo "== big on (default)"; python profiles/quickbench.py
echo "== big off"; MRP_BIG=0 python profiles/quickbench.py
echo "== big on, solver ctas 3"; MRP_SOLVER_CTAS=3 python profiles/quickbench.py
echo "== 262144 big on / off"; QB_ENVS=262144 python profiles/quickbench.py; MRP_BIG=0 QB_ENVS=262144 python profiles/quickbench.py
echo "== 524288 big on / off"; QB_ENVS=524288 python profiles/quickbench.py; MRP_BIG=0 QB_ENVS=524288 python profiles/quickbench.py
ALL=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
