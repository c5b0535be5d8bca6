# wave split sweep for mrp_step_host (back chunks in the first wave, of 8)
for S in 2 3 4; do echo "== MRP_HOST_WAVE_SPLIT=$S"; MRP_HOST_WAVE_SPLIT=$S QB_E2E=1 python profiles/quickbench.py; MRP_HOST_WAVE_SPLIT=$S QB_E2E=1 python profiles/quickbench.py MultiRobotPuzzle-v0 MultiRobotPuzzle-v2; MRP_HOST_WAVE_SPLIT=$S QB_ENVS=524288 QB_E2E=1 python profiles/quickbench.py; done
MRP_HOST_WAVE_SPLIT=3 MRP_TRACE=1 QB_E2E=1 python profiles/quickbench.py 2>&1 | grep "h2d_done" | tail -2
