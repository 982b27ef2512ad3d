# A/B of k_hour variants (exp/lib_*.so built by tools/exp.py): parity first for each, then bench at 10 M and 1 M, twice
mkdir -p gpurun_out
: > gpurun_out/r2t.txt
for v in $VARIANTS; do
  EPI_LIB=$PWD/exp/lib_$v.so timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2t_pytest_$v.log 2>&1; echo "$v pytest rc=$? $(tail -1 gpurun_out/r2t_pytest_$v.log)" >> gpurun_out/r2t.txt
done
for rep in 1 2; do
for v in $VARIANTS; do
for wl in 10m 1m; do
  EPI_LIB=$PWD/exp/lib_$v.so timeout 300 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/r2t_${v}_$wl.json 2> gpurun_out/r2t_${v}_$wl.err
  python - <<PY >> gpurun_out/r2t.txt
import json
try:
    j = json.loads(open("gpurun_out/r2t_${v}_$wl.json").read().strip().splitlines()[-1])
    print("%-10s %-4s value %.4e ms/day %.4f frac %.4f" % ("$v", "$wl", j["value"], j["ms_per_step"], j["roofline"]["frac"]), {k: round(x, 4) for k, x in j["roofline"]["per_kernel_ms"].items() if k.startswith("k_")}, j["config"]["last_counts_row"])
except Exception as ex:
    print("$v $wl FAILED", ex)
PY
done; done; done
cat gpurun_out/r2t.txt
