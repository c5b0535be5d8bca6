python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
b() { python bench.py --config $1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 6 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 value %.3e ms %.3f e2e ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['episode_stats']['episodes'])"; }
echo "== spares off"; MRP_SPARES=0 b c3-resets; MRP_SPARES=0 b c3
echo "== spares on (default)"; b c3-resets; b c3
MRP_SPARES=0 QB_ENVS=524288 python profiles/quickbench.py; QB_ENVS=524288 python profiles/quickbench.py
MRP_SPARES=0 QB_ENVS=262144 python profiles/quickbench.py; QB_ENVS=262144 python profiles/quickbench.py
MRP_SPARES=0 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2;  python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2
