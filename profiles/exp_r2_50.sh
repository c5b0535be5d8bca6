# main chain of mrp_step on a library stream of middle priority (MRP_MAIN_PRIO=1) vs on the caller's stream
# (the code under test was measured and NOT kept: DESIGN.md §8, "Measured dead ends"; this script is the record of the A/B)
for i in 1 2; do echo "== MRP_MAIN_PRIO=0"; python profiles/quickbench.py; echo "== MRP_MAIN_PRIO=1"; MRP_MAIN_PRIO=1 python profiles/quickbench.py; done
echo "== trace 0 / 1"
MRP_TRACE=1 python profiles/quickbench.py | grep "mrp trace" | tail -2
MRP_MAIN_PRIO=1 MRP_TRACE=1 python profiles/quickbench.py | grep "mrp trace" | tail -2
echo "== 262144 / 524288 / v0 : 0, 1"
QB_ENVS=262144 python profiles/quickbench.py; MRP_MAIN_PRIO=1 QB_ENVS=262144 python profiles/quickbench.py
QB_ENVS=524288 python profiles/quickbench.py; MRP_MAIN_PRIO=1 QB_ENVS=524288 python profiles/quickbench.py
python profiles/quickbench.py MultiRobotPuzzle-v0; MRP_MAIN_PRIO=1 python profiles/quickbench.py MultiRobotPuzzle-v0
