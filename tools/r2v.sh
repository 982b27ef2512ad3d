# per-hour-of-day kernel times of k_hour variants
mkdir -p gpurun_out
: > gpurun_out/r2v.txt
for v in $VARIANTS; do
  EPI_LIB=$PWD/exp/lib_$v.so timeout 300 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --workload 10m > gpurun_out/r2v_${v}.json 2> gpurun_out/r2v_${v}.err
  python - <<PY >> gpurun_out/r2v.txt
import json
try:
    j = json.loads(open("gpurun_out/r2v_${v}.json").read().strip().splitlines()[-1])
    ph = j["roofline"]["per_hour_of_day_ms"]
    print("%-10s ms/day %.4f frac %.4f" % ("$v", j["ms_per_step"], j["roofline"]["frac"]), " ".join("h%s %.0f+%.0f" % (h, ph[h]["hour_ms"]*1e3, ph[h]["commit_ms"]*1e3) for h in ("0","7","8","9","12","16","17","20","23")))
except Exception as ex:
    print("$v FAILED", ex)
PY
done
cat gpurun_out/r2v.txt
