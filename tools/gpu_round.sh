#!/bin/bash
# One GPU session for the record (profiles/): parity tests, smoke, bench lines (ours + reference arm), ncu launch list of a whole
# simulated day with DRAM bytes, full ncu captures of k_hour (work hour, home hour) and k_commit.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
timeout 300 python bench.py --workload 1m --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour -s 40 -c 1 -o gpurun_out/prof_hour_work python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hour_work.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour -s 49 -c 1 -o gpurun_out/prof_hour_home python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hour_home.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_commit -s 40 -c 1 -o gpurun_out/prof_commit_work python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_commit_work.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_commit -s 49 -c 1 -o gpurun_out/prof_commit_home python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_commit_home.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_10m.json; tail -3 gpurun_out/bench_10m.err; cat gpurun_out/bench_1m.json; cat gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
