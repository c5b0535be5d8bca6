python profiles/diag_square.py
MRP_STAGE_ROWS=0 MRP_GRAPH=0 python profiles/diag_square.py
python profiles/diag_square.py MultiRobotPuzzleHeavy-v2
python profiles/diag_square.py MultiRobotPuzzleHeavy-v0
