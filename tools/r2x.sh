# Final record session, round 2 (one GPU): -m gpu suite, smoke, bench lines (config #3 default / as written = 1080 h / the 10 days the
# multi-GPU lines time, config #2 default / as written, reference arm), ncu launch list + full captures of k_hour (work, home) and k_commit_lanes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2x_gpu.txt 2>&1; nproc >> gpurun_out/r2x_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2x_smoke.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2x_bench_10m.json 2> gpurun_out/r2x_bench_10m.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench_10m_days4to13.json 2> gpurun_out/r2x_bench_10m_days4to13.err
timeout 300 python bench.py --steps 42 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench_cfg3_1080h.json 2> gpurun_out/r2x_bench_cfg3_1080h.err
timeout 300 python bench.py --workload 1m --steps 42 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench_cfg2_1080h.json 2> gpurun_out/r2x_bench_cfg2_1080h.err
timeout 300 python bench.py --workload 1m --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench_1m.json 2> gpurun_out/r2x_bench_1m.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour -s 40 -c 1 -o gpurun_out/r2x_prof_hour_work python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_ncu_hour_work.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hour -s 49 -c 1 -o gpurun_out/r2x_prof_hour_home python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_ncu_hour_home.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_commit -s 40 -c 1 -o gpurun_out/r2x_prof_commit_work python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_ncu_commit_work.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_commit -s 49 -c 1 -o gpurun_out/r2x_prof_commit_home python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_ncu_commit_home.log 2>&1
tail -3 gpurun_out/r2x_pytest_gpu.log; tail -2 gpurun_out/r2x_smoke.log
for f in 10m 10m_days4to13 cfg3_1080h cfg2_1080h 1m; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2x_bench_$f.json') if l.startswith('{')][-1])
    print('$f', '%.4g'%d['value'], 'ms/day %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'mov frac %.4f'%d['roofline']['frac'], 'day frac %.4f'%d['roofline']['whole_day']['frac'], d['config']['phase'].get('ms_per_day_open'), d['config']['phase'].get('ms_per_day_locked_down'))
except Exception as ex: print('$f FAILED', ex)
PY
done
cat gpurun_out/r2x_bench_reference.json | cut -c1-300
