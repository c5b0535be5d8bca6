echo "== default"; QB_E2E=1 python profiles/quickbench.py
for k in 2 4 8; do echo "== host pipe $k"; MRP_HOST_PIPE=$k QB_E2E=1 python profiles/quickbench.py; done
