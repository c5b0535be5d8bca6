python -m pytest tests/test_gpu_parity.py tests/test_identical_states.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
QB_PHASES=1 python profiles/quickbench.py
python profiles/quickbench.py
MRP_OVERLAP_POST=0 python profiles/quickbench.py
QB_ENVS=262144 python profiles/quickbench.py
QB_PHASES=1 QB_ENVS=262144 python profiles/quickbench.py
ALL=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
