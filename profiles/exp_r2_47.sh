# k_pre with 8 shared-memory words per body (origin recomputed, rotation = q): 72 instead of 102 words per lane, 5 resident CTAs
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
OLD=gym_puzzles_b200/csrc/build/var/libmrp_old.so
for i in 1 2; do echo "== old (13 words)"; MRP_LIB_PATH=$OLD python profiles/quickbench.py; echo "== new (8 words)"; python profiles/quickbench.py; done
echo "== phases old / new"; MRP_LIB_PATH=$OLD QB_PHASES=1 python profiles/quickbench.py; QB_PHASES=1 python profiles/quickbench.py
echo "== 524288 / v0 / v2 old, new"
MRP_LIB_PATH=$OLD QB_ENVS=524288 python profiles/quickbench.py; QB_ENVS=524288 python profiles/quickbench.py
MRP_LIB_PATH=$OLD python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2; python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
