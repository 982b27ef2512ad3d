mkdir -p gpurun_out
python tools/hours.py 10m 3 3 > gpurun_out/r2d_hours_tiles.txt 2>&1; cat gpurun_out/r2d_hours_tiles.txt
EPI_TILES=0 python tools/hours.py 10m 3 3 > gpurun_out/r2d_hours_notiles.txt 2>&1; cat gpurun_out/r2d_hours_notiles.txt
EPI_TILE_OFFICES=32 EPI_TILE_HOUSES=256 EPI_TILE_HOUSE_THREADS=256 python tools/hours.py 10m 3 3 > gpurun_out/r2d_hours_big.txt 2>&1; cat gpurun_out/r2d_hours_big.txt
EPI_TILE_OFFICES=8 EPI_TILE_HOUSES=64 EPI_TILE_HOUSE_THREADS=128 EPI_TILE_OFFICE_THREADS=128 python tools/hours.py 10m 3 3 > gpurun_out/r2d_hours_small.txt 2>&1; cat gpurun_out/r2d_hours_small.txt
