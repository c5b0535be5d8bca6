python -m pytest tests/test_gpu_parity.py tests/test_identical_states.py tests/test_golden.py tests/test_gpu_vector_env.py tests/test_next_rows.py -m gpu -x -q 2>&1 | tail -3
python profiles/quickbench.py; QB_PHASES=1 python profiles/quickbench.py
QB_ENVS=262144 python profiles/quickbench.py
ALL=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
QB_AGENTS=5 QB_ENVS=262144 python profiles/quickbench.py MultiRobotPuzzle-v2
