# A/B of programmatic dependent launch (EPI_PDL=1) on config #2 (1 M, L2-resident) and config #3 (10 M), parity first
mkdir -p gpurun_out
EPI_PDL=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_travel_gpu.py -m gpu -x -q > gpurun_out/r2r_pytest_pdl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest_pdl.log
tail -3 gpurun_out/r2r_pytest_pdl.log
for pdl in 0 1 0 1; do
for wl in 1m 10m; do
EPI_PDL=$pdl timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_${wl}_pdl$pdl.json 2> gpurun_out/r2r_bench_${wl}_pdl$pdl.err
python - <<PY
import json
txt=open('gpurun_out/r2r_bench_${wl}_pdl$pdl.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if line:
    d=json.loads(line[-1]); print('pdl=$pdl $wl', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['roofline']['frac'])
else: print('pdl=$pdl $wl FAILED'); print(open('gpurun_out/r2r_bench_${wl}_pdl$pdl.err').read()[-800:])
PY
done; done
