# weak-scaling lines with the final kernels (8-GPU box): N = 8, 4, 2 over the same ten days
mkdir -p gpurun_out
run() {
  name=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((10 + RANDOM % 80)) bench.py --gpus $n "$@" > gpurun_out/r2y_$name.json 2> gpurun_out/r2y_$name.err
  python - <<PY
import json
txt=open('gpurun_out/r2y_$name.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if line:
    d=json.loads(line[-1]); print('$name', '%.4g'%d['value'], 'ms/day %.4f'%d['ms_per_step'], 'launches', d['gpu_launches'], 'travel ms/day', d['roofline']['per_kernel_ms']['travel_kernels_per_day_ms'], 'e2e %.4g'%d['e2e']['value'])
else:
    print('$name FAILED'); print(open('gpurun_out/r2y_$name.err').read()[-1500:])
PY
}
run n8_10m 8 --workload 10m --steps 10 --warmup 3
run n4_10m 4 --workload 10m --steps 10 --warmup 3
run n2_10m 2 --workload 10m --steps 10 --warmup 3
