mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_travel_gpu.py tests/test_engine_app_gpu.py -m gpu -x -q --durations=5 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -30 gpurun_out/r2j_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload 10m --steps 10 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_n2.json')); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['per_kernel_ms'], d['config']['phase'])"; tail -5 gpurun_out/r2j_bench_n2.err
