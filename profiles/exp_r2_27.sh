# spare episodes: parity, then the steady-reset bench (TimeLimit 200) with and without, lanes per warp of the refill kernel
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_next_rows.py -m gpu -x -q 2>&1 | tail -2
b() { python bench.py --config $1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 6 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 value %.3e ms %.3f e2e ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['episode_stats']['episodes'])"; }
echo "== spares off"; MRP_SPARES=0 b c3-resets
for L in 32 8 4; do echo "== spares on, lanes $L"; MRP_RESET_LANES=$L b c3-resets; done
echo "== default config"; MRP_SPARES=0 b c3; b c3
echo "== c2 / c4"; b c2; b c4
