echo "== baseline old flow"; QB_PHASES=1 python profiles/quickbench.py
python profiles/quickbench.py
for c in 75 50 25; do echo "== carveout pre $c"; MRP_CARVEOUT_PRE=$c QB_PHASES=1 python profiles/quickbench.py; done
for c in 75 50 25; do echo "== carveout post $c"; MRP_CARVEOUT_POST=$c QB_PHASES=1 python profiles/quickbench.py; done
for c in 75 50; do echo "== carveout broad $c"; MRP_CARVEOUT_BROAD=$c QB_PHASES=1 python profiles/quickbench.py; done
echo "== front flow carve 50"; MRP_FRONT=1 MRP_CARVEOUT_POST=50 QB_PHASES=1 python profiles/quickbench.py
