# front-half waves of mrp_step_host: 1 / 2 / 3 / 4
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for W in 1 2 3 4; do echo "== MRP_HOST_WAVES=$W"; MRP_HOST_WAVES=$W QB_E2E=1 python profiles/quickbench.py; MRP_HOST_WAVES=$W QB_E2E=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2; MRP_HOST_WAVES=$W QB_ENVS=524288 QB_E2E=1 python profiles/quickbench.py;  MRP_HOST_WAVES=$W QB_ENVS=262144 QB_E2E=1 python profiles/quickbench.py; done
MRP_HOST_WAVES=3 MRP_TRACE=1 QB_E2E=1 python profiles/quickbench.py 2>&1 | grep "h2d_done" | tail -2
MRP_HOST_WAVES=4 MRP_TRACE=1 QB_E2E=1 python profiles/quickbench.py 2>&1 | grep "h2d_done" | tail -2
