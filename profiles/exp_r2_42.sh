# persistent-solver knobs (compile time): lanes idle before a refill, trips between refill checks
V=gym_puzzles_b200/csrc/build/var
echo "== default (refill 8, trips 8)"; python profiles/quickbench.py; QB_PHASES=1 python profiles/quickbench.py
for v in r4 r16 t4 t16 r16t16; do echo "== $v"; MRP_LIB_PATH=$V/libmrp_$v.so python profiles/quickbench.py; MRP_LIB_PATH=$V/libmrp_$v.so QB_PHASES=1 python profiles/quickbench.py; done
