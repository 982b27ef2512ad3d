mkdir -p gpurun_out
# launch list of one day with dram bytes (tiles on)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/r2e_launches.csv python tools/hours.py 10m 1 2 > gpurun_out/r2e_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour_tile -s 8 -c 1 -o gpurun_out/r2e_tile_office python tools/hours.py 10m 1 2 > gpurun_out/r2e_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour_tile -s 20 -c 1 -o gpurun_out/r2e_tile_house python tools/hours.py 10m 1 2 > gpurun_out/r2e_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour_list -s 8 -c 1 -o gpurun_out/r2e_list python tools/hours.py 10m 1 2 > gpurun_out/r2e_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
