python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python profiles/quickbench.py
MRP_OVERLAP_POST=0 python profiles/quickbench.py
QB_ENVS=524288 python profiles/quickbench.py
QB_ENVS=262144 python profiles/quickbench.py
QB_E2E=1 python profiles/quickbench.py MultiRobotPuzzle-v0
