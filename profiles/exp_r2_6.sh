echo "== default"; python profiles/quickbench.py; QB_PHASES=1 python profiles/quickbench.py
echo "== events in k_post"; MRP_EVENTS_IN_POST=1 python profiles/quickbench.py; MRP_EVENTS_IN_POST=1 QB_PHASES=1 python profiles/quickbench.py
echo "== 262144"; QB_ENVS=262144 python profiles/quickbench.py; MRP_EVENTS_IN_POST=1 QB_ENVS=262144 python profiles/quickbench.py
MRP_EVENTS_IN_POST=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
