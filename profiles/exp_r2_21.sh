# k_solve_big in the front half of mrp_step_host: parity (chunked host path), end-to-end A/B, timeline
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for B in 0 1; do echo "== MRP_BIG=$B"; MRP_BIG=$B QB_E2E=1 python profiles/quickbench.py; done
MRP_TRACE=1 QB_E2E=1 python profiles/quickbench.py 2>&1 | grep "h2d_done" | tail -3
