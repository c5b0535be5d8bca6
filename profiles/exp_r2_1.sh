for c in 1 2 3 4; do MRP_SOLVER_CTAS=$c python profiles/quickbench.py; done
MRP_SOLVER_CTAS=2 MRP_CHUNKS=2 python profiles/quickbench.py
MRP_SOLVER_CTAS=1 MRP_CHUNKS=2 python profiles/quickbench.py
MRP_SOLVER_CTAS=2 MRP_CHUNKS=4 python profiles/quickbench.py
QB_PHASES=1 QB_ENVS=262144 python profiles/quickbench.py
QB_PHASES=1 QB_ENVS=2097152 python profiles/quickbench.py
MRP_FUSED_STEP=1 python profiles/small_batch.py MultiRobotPuzzleHeavy-v0 2>&1 | head -4
MRP_FUSED_STEP=1 python profiles/small_batch.py MultiRobotPuzzle-v0 2>&1 | head -4
python profiles/small_batch.py MultiRobotPuzzleHeavy-v0 2>&1 | head -3
