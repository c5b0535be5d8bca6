for D in 0 1 2 3; do echo "== MRP_REFILL_DEBUG=$D"; MRP_REFILL_DEBUG=$D python profiles/quickbench.py; done
MRP_SPARES=0 python profiles/quickbench.py
MRP_TRACE=1 python profiles/quickbench.py 2>&1 | grep "mrp trace" | tail -3
MRP_SPARES=0 MRP_TRACE=1 python profiles/quickbench.py 2>&1 | grep "mrp trace" | tail -3
