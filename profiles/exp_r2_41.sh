for c in c2 c5; do python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_${c}.json 2> gpurun_out/r2c_bench_${c}.err; done
python -m pytest tests/test_square_variant.py tests/test_gpu_vector_env.py -m gpu -q 2>&1 | tail -2
