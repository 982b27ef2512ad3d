# 8-GPU record session: weak-scaling line (8 x config #3 regions), BASELINE config #4 (8 x 2 M) and #5 (8 x 20 M, 2160 h) as written,
# the 2-rank CLI / travel parity tests, GPU timeline of the exchanges at N = 8.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/r2s_gpus.txt 2>&1
run() {  # name, gpus, extra args
  name=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((10 + RANDOM % 80)) bench.py --gpus $n "$@" > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err
  python - <<PY
import json
txt=open('gpurun_out/r2s_$name.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if line:
    d=json.loads(line[-1]); print('$name', '%.4g'%d['value'], 'ms/day %.4f'%d['ms_per_step'], 'launches', d['gpu_launches'], 'travel ms/day', d['roofline']['per_kernel_ms']['travel_kernels_per_day_ms'], 'e2e %.4g'%d['e2e']['value'])
else:
    print('$name FAILED'); print(open('gpurun_out/r2s_$name.err').read()[-1500:])
PY
}
run n8_10m 8 --workload 10m --steps 10 --warmup 3
run n8_cfg4_2m 8 --workload 2m --steps 30 --warmup 3
run n8_cfg5_20m 8 --workload 20m --steps 87 --warmup 3
run n4_10m 4 --workload 10m --steps 10 --warmup 3
run n2_10m 2 --workload 10m --steps 10 --warmup 3
EPI_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 tools/timeline.py 10m 2 > gpurun_out/r2s_tl.log 2>&1
cp gpurun_out/timeline_r0.txt gpurun_out/r2s_timeline_n8_r0.txt; cp gpurun_out/timeline_r7.txt gpurun_out/r2s_timeline_n8_r7.txt
grep -E "leave|arriv" gpurun_out/r2s_timeline_n8_r7.txt | head -12
timeout 600 python -m pytest tests/test_engine_app_gpu.py tests/test_travel_gpu.py -m gpu -x -q > gpurun_out/r2s_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest_multi.log
tail -3 gpurun_out/r2s_pytest_multi.log
