# CUDA-graph replay of small-batch steps: parity, then latency with and without (MRP_GRAPH), fused single kernel for comparison
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_next_rows.py tests/test_error_paths.py -m gpu -x -q 2>&1 | tail -3
for G in 0 1; do echo "== MRP_GRAPH=$G"; MRP_GRAPH=$G SB_SIZES=1,6,64,1024,16384 python profiles/small_batch.py MultiRobotPuzzleHeavy-v0; MRP_GRAPH=$G SB_SIZES=1,6,64 python profiles/small_batch.py MultiRobotPuzzle-v0; done
echo "== fused single kernel (MRP_FUSED_STEP=1, no graph)"; MRP_FUSED_STEP=1 SB_SIZES=1,6,64 python profiles/small_batch.py MultiRobotPuzzleHeavy-v0
