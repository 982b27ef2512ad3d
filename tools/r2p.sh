mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload 10m --steps 10 --warmup 3 > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err
for n in 1 2; do python - <<PY
import json
txt=open('gpurun_out/r2p_bench_n$n.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
d=json.loads(line[-1]); print('N=$n', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['config']['phase'])
PY
done
tail -3 gpurun_out/r2p_bench_n1.err
