mkdir -p gpurun_out
for v in base nonccl; do
  if [ $v = nonccl ]; then export EPI_DEBUG_NO_NCCL=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload 10m --steps 10 --warmup 3 > gpurun_out/r2k_$v.json 2> gpurun_out/r2k_$v.err
  python - <<PY
import json
txt=open('gpurun_out/r2k_$v.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
d=json.loads(line[-1]); print('$v', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['per_kernel_ms']['travel_kernels_per_day_ms'])
PY
done
