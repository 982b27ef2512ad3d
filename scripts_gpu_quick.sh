#!/bin/bash
# quick GPU check: parity tests + bench lines + ncu launch list (+ full ncu capture of k_hour / k_commit when NCU=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
timeout 300 python bench.py --workload 1m --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
if [ "$NCU" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hour.1 -s 40 -c 5 -o gpurun_out/prof_hour python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hour.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_commit -s 48 -c 2 -o gpurun_out/prof_commit python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_commit.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_10m.json; tail -3 gpurun_out/bench_10m.err; cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err
