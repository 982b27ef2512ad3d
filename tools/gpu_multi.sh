#!/bin/bash
# multi-GPU session: all gpu tests (the 2-region CLI test needs 2 GPUs) + the N-GPU bench lines
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_multi.log
tail -15 gpurun_out/pytest_gpu_multi.log
for n in ${NS:-2}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload ${WL:-2m} --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
tail -c 1500 gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
done
