mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2b_gpus.txt
timeout 900 python -m pytest tests/test_engine_app_gpu.py tests/test_travel_gpu.py tests/test_parity_gpu.py -m gpu -x -q --durations=8 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload 10m --steps 10 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
tail -c 2500 gpurun_out/r2b_bench_n2.json; tail -5 gpurun_out/r2b_bench_n2.err
