# step timeline at configs[1] (v0, 65,536 envs) and at 262,144 Heavy-v0 envs
MRP_TRACE=1 QB_ENVS=65536 python profiles/quickbench.py MultiRobotPuzzle-v0 | grep -E "mrp trace|ms/step" | tail -4
MRP_TRACE=1 QB_ENVS=262144 python profiles/quickbench.py | grep -E "mrp trace|ms/step" | tail -3
for v in "MRP_BIG=0" "MRP_OVERLAP_POST=0" "MRP_BIG=0 MRP_OVERLAP_POST=0" "MRP_GRAPH=1 MRP_BIG=0 MRP_OVERLAP_POST=0"; do echo "== $v"; env $v QB_ENVS=65536 python profiles/quickbench.py MultiRobotPuzzle-v0; done
echo "== default"; QB_ENVS=65536 python profiles/quickbench.py MultiRobotPuzzle-v0
