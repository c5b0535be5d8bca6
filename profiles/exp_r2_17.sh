python -m pytest tests/test_square_variant.py -m gpu -q 2>&1 | grep -v "^$" | tail -40
for B in 0 1; do echo "== MRP_BIG=$B"; MRP_BIG=$B QB_ENVS=524288 python profiles/quickbench.py MultiRobotPuzzleSquare-v2; done
MRP_BIG=0 MRP_OVERLAP_POST=1 QB_ENVS=524288 python profiles/quickbench.py MultiRobotPuzzleSquare-v2
