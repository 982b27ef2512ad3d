#!/bin/bash
# GPU session r01d: parity tests, A/B of the agent numbering / PLAIN-hour kernel / occupancy target, bench lines, ncu day view + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/exp.txt
ab() {  # name order lib
  EPI_AGENT_ORDER=$2 EPI_LIB=$PWD/exp/lib_$3.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/exp_$1.json 2> gpurun_out/exp_$1.err
  python - <<PY >> gpurun_out/exp.txt
import json
try:
    j = json.loads(open("gpurun_out/exp_$1.json").read().strip().splitlines()[-1])
    print("$1 value %.4e e2e %.4e ms/day %.3f frac %.3f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["frac"]), {k: round(x, 4) for k, x in j["roofline"]["per_kernel_ms"].items()}, j["clocks"], j["config"]["last_counts_row"])
except Exception as ex:
    print("$1 FAILED", ex)
PY
}
ab creation_plainoff creation plainoff
ab house_plainoff house plainoff
ab house_base house base
ab house_minb8 house minb8
ab creation_base creation base
cat gpurun_out/exp.txt
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
timeout 300 python bench.py --workload 1m --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
cat gpurun_out/bench_10m.json; tail -3 gpurun_out/bench_10m.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour -s 40 -c 1 -o gpurun_out/prof_hour_work python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hour_work.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour -s 49 -c 1 -o gpurun_out/prof_hour_home python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hour_home.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_commit -s 40 -c 1 -o gpurun_out/prof_commit_work python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_commit_work.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_commit -s 49 -c 1 -o gpurun_out/prof_commit_home python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_commit_home.log 2>&1
ls -la gpurun_out
