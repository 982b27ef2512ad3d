mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_travel_gpu.py tests/test_engine_app_gpu.py -m gpu -x -q --durations=6 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -12 gpurun_out/r2o_pytest.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --workload 10m --steps 10 --warmup 3 > gpurun_out/r2o_bench_n$n.json 2> gpurun_out/r2o_bench_n$n.err
python - <<PY
import json
txt=open('gpurun_out/r2o_bench_n$n.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
d=json.loads(line[-1]); print('N=$n', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['per_kernel_ms']['travel_kernels_per_day_ms'])
PY
done
