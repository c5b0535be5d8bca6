# action rows in / observation rows out through shared memory (coalesced float4 runs): parity, then A/B (MRP_STAGE_ROWS)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for S in 0 1; do echo "== MRP_STAGE_ROWS=$S"; MRP_STAGE_ROWS=$S python profiles/quickbench.py; MRP_STAGE_ROWS=$S QB_PHASES=1 python profiles/quickbench.py;  MRP_STAGE_ROWS=$S QB_E2E=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2; done
for S in 0 1; do echo "== MRP_STAGE_ROWS=$S"; MRP_STAGE_ROWS=$S python profiles/quickbench.py; done
