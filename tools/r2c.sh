mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiles_gpu.py tests/test_parity_gpu.py -m gpu -x -q --durations=8 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -30 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 1800 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
EPI_TILES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_notiles.json 2> gpurun_out/r2c_bench_notiles.err
tail -c 1800 gpurun_out/r2c_bench_notiles.json; tail -5 gpurun_out/r2c_bench_notiles.err
