mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tiles_gpu.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -5 gpurun_out/r2h_pytest.log
python tools/hours.py 10m 3 3 > gpurun_out/r2h_hours_tiles.txt 2>&1; cat gpurun_out/r2h_hours_tiles.txt
EPI_TILE_OFFICES=16 EPI_TILE_HOUSES=128 python tools/hours.py 10m 3 3 > gpurun_out/r2h_hours_big.txt 2>&1; cat gpurun_out/r2h_hours_big.txt
