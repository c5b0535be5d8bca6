# final verification of the round-2 tree: full GPU suite, smoke, both bench arms
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; tail -c 400 gpurun_out/r2e_bench_ref.json
python bench.py > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_default.err; tail -c 300 gpurun_out/r2e_bench_default.json
