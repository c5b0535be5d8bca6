# mrp_step_host: small copies ahead of the bulk copy, wave split
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for S in 4 5 6 3; do echo "== MRP_HOST_WAVE_SPLIT=$S"; MRP_HOST_WAVE_SPLIT=$S QB_E2E=1 python profiles/quickbench.py; done
MRP_HOST_WAVES=1 QB_E2E=1 python profiles/quickbench.py
MRP_TRACE=1 QB_E2E=1 python profiles/quickbench.py 2>&1 | grep "h2d_done" | tail -2
MRP_HOST_WAVE_SPLIT=5 MRP_TRACE=1 QB_E2E=1 python profiles/quickbench.py 2>&1 | grep "h2d_done" | tail -2
