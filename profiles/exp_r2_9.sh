echo "== timed, split on"; QB_PHASES=1 python profiles/quickbench.py
echo "== timed, split off"; MRP_VEL_SPLIT=0 QB_PHASES=1 python profiles/quickbench.py
export MRP_LIB_PATH=$PWD/gym_puzzles_b200/csrc/libmrp_probe.so
MRP_BIG=0 python profiles/tailprobe.py 2>&1 | tail -9 | head -7
