# invMass / invI / localCenter from the per-CTA constant table (one shared-memory read) instead of select chains over kernel parameters
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_square_variant.py -m gpu -x -q 2>&1 | tail -2
OLD=gym_puzzles_b200/csrc/build/var/libmrp_old.so
for i in 1 2 3; do echo "== old"; MRP_LIB_PATH=$OLD python profiles/quickbench.py; echo "== new"; python profiles/quickbench.py; done
echo "== phases old / new"; MRP_LIB_PATH=$OLD QB_PHASES=1 python profiles/quickbench.py; QB_PHASES=1 python profiles/quickbench.py
echo "== v0 / v2 old, new"
MRP_LIB_PATH=$OLD python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2; python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
