# k_solve_vel compiled for 5 / 6 resident CTAs per SM (96 / 80 registers instead of 120) with MRP_SOLVER_CTAS = 5 / 6: more chains in flight
# (the code under test was measured and NOT kept: DESIGN.md §8, "Measured dead ends"; this script is the record of the A/B)
V=gym_puzzles_b200/csrc/build/var
for i in 1 2; do
echo "== default (4 CTAs)"; python profiles/quickbench.py
echo "== 5 CTAs"; MRP_LIB_PATH=$V/libmrp_lb5.so MRP_SOLVER_CTAS=5 python profiles/quickbench.py
echo "== 6 CTAs"; MRP_LIB_PATH=$V/libmrp_lb6.so MRP_SOLVER_CTAS=6 python profiles/quickbench.py
echo "== default lib, 5 CTAs requested"; MRP_SOLVER_CTAS=5 python profiles/quickbench.py
done
echo "== 262144: default / 5 / 6"; QB_ENVS=262144 python profiles/quickbench.py; MRP_LIB_PATH=$V/libmrp_lb5.so MRP_SOLVER_CTAS=5 QB_ENVS=262144 python profiles/quickbench.py; MRP_LIB_PATH=$V/libmrp_lb6.so MRP_SOLVER_CTAS=6 QB_ENVS=262144 python profiles/quickbench.py
