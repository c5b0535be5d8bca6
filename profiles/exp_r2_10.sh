for m in 0 2; do echo "== vel mode $m"; MRP_VEL_SPLIT=$m QB_PHASES=1 python profiles/quickbench.py; MRP_VEL_SPLIT=$m python profiles/quickbench.py; MRP_VEL_SPLIT=$m QB_ENVS=262144 python profiles/quickbench.py; done
MRP_VEL_SPLIT=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout" 2>&1 | tail -2
