mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_travel_gpu.py tests/test_engine_app_gpu.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
EPI_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/timeline.py 10m 2 > gpurun_out/r2m_tl.log 2>&1
grep -E "leave|arriv" gpurun_out/timeline_r1.txt | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload 10m --steps 10 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err
python - <<PY
import json
txt=open('gpurun_out/r2m_bench_n2.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
d=json.loads(line[-1]); print('N=2', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['per_kernel_ms']['travel_kernels_per_day_ms'])
PY
