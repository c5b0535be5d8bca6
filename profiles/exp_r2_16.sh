# square variant (configs[4]) on the device: parity, throughput; whole GPU suite
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
QB_ENVS=524288 python profiles/quickbench.py MultiRobotPuzzleSquare-v2
QB_ENVS=524288 QB_PHASES=1 python profiles/quickbench.py MultiRobotPuzzleSquare-v2
MRP_OVERLAP_POST=1 QB_ENVS=524288 python profiles/quickbench.py MultiRobotPuzzleSquare-v2
QB_ENVS=1048576 python profiles/quickbench.py MultiRobotPuzzleSquare-v2
python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
