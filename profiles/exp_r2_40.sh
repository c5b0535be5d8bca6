# round-2 final evidence: GPU suite, smoke, bench lines of every config, reference arm, ncu launch list + full capture, small batches, checking build
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/r2c_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> gpurun_out/r2c_gputest.log
python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
python bench.py --impl reference > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err
bash profiles/run_configs_1gpu.sh r2c > gpurun_out/r2c_configs.log 2>&1
BENCH_LIST=1 bash profiles/capture.sh r2c > gpurun_out/r2c_capture.log 2>&1
SB_SIZES=1,6,64,1024 python profiles/small_batch.py MultiRobotPuzzleHeavy-v0 > gpurun_out/r2c_small_batch.log 2>&1
