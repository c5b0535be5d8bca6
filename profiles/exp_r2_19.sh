python -m pytest tests/test_square_variant.py -m gpu -q 2>&1 | tail -3
python profiles/diag_square.py
QB_ENVS=524288 python profiles/quickbench.py MultiRobotPuzzleSquare-v2
QB_ENVS=524288 QB_PHASES=1 python profiles/quickbench.py MultiRobotPuzzleSquare-v2
python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
