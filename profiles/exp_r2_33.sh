for D in 0 1 2 3 4 5 6 7 8 9; do echo "== pass stops after $D kernels"; MRP_REFILL_DEBUG=$D python profiles/quickbench.py; done
