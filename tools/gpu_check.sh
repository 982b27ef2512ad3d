#!/bin/bash
# parity tests + bench lines (+ optional extra command in $EXTRA)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
timeout 300 python bench.py --workload 1m --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
python tools/summarize.py gpurun_out | head -3; tail -2 gpurun_out/bench_10m.err
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
