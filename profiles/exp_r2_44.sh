# checking build over every kernel family incl. waves / spares / square variant; k_solve_big CTAs per SM
MRP_LIB_PATH=$PWD/gym_puzzles_b200/csrc/libmrp_check.so python profiles/check_run.py 2>&1 | tail -4
for B in 1 2; do echo "== MRP_BIG_CTAS=$B"; MRP_BIG_CTAS=$B python profiles/quickbench.py; MRP_BIG_CTAS=$B QB_ENVS=524288 python profiles/quickbench.py; done
b() { python bench.py --config $1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 4 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 value %.3e ms %.3f e2e ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"; }
MRP_BIG_CTAS=2 b c3-resets
