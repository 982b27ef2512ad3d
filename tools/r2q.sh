# Round-2 record session (one GPU): the whole -m gpu suite, smoke, bench lines (config #3, #2, reference arm), ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2q_smoke.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2q_bench_10m.json 2> gpurun_out/r2q_bench_10m.err
timeout 300 python bench.py --workload 1m --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_1m.json 2> gpurun_out/r2q_bench_1m.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_ncu_launches.log 2>&1
tail -4 gpurun_out/r2q_pytest_gpu.log; tail -2 gpurun_out/r2q_smoke.log; cat gpurun_out/r2q_bench_10m.json; tail -3 gpurun_out/r2q_bench_10m.err; cat gpurun_out/r2q_bench_1m.json; cat gpurun_out/r2q_bench_reference.json
