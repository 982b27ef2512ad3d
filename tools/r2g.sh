mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour_tile -s 8 -c 1 -o gpurun_out/r2g_tile_house python tools/hours.py 10m 1 2 > gpurun_out/r2g_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour_tile -s 3 -c 1 -o gpurun_out/r2g_tile_office python tools/hours.py 10m 1 2 > gpurun_out/r2g_b.log 2>&1
ls -la gpurun_out/r2g*.ncu-rep
