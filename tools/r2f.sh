mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiles_gpu.py tests/test_parity_gpu.py -m gpu -x -q --durations=5 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -30 gpurun_out/r2f_pytest.log
python tools/hours.py 10m 3 3 > gpurun_out/r2f_hours_tiles.txt 2>&1; cat gpurun_out/r2f_hours_tiles.txt
EPI_TILE_OFFICES=4 EPI_TILE_HOUSES=32 python tools/hours.py 10m 3 3 > gpurun_out/r2f_hours_small.txt 2>&1; cat gpurun_out/r2f_hours_small.txt
EPI_TILE_OFFICES=16 EPI_TILE_HOUSES=128 python tools/hours.py 10m 3 3 > gpurun_out/r2f_hours_big.txt 2>&1; cat gpurun_out/r2f_hours_big.txt
