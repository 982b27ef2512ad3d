#!/bin/bash
# A/B with the agent numbering as a second axis: $RUNS = "name:order:lib ..."
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -4 gpurun_out/pytest_gpu.log
fi
: > gpurun_out/exp.txt
for run in ${RUNS}; do
  IFS=: read name order lib <<< "$run"
  EPI_AGENT_ORDER=$order EPI_LIB=$PWD/exp/lib_$lib.so timeout 300 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --workload ${WL:-10m} > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  python - <<PY >> gpurun_out/exp.txt
import json
try:
    j = json.loads(open("gpurun_out/exp_$name.json").read().strip().splitlines()[-1])
    print("%-18s value %.4e e2e %.4e ms/day %.3f frac %.3f" % ("$name", j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["frac"]), {k: round(x, 4) for k, x in j["roofline"]["per_kernel_ms"].items() if k != "travel_kernels_total_ms"}, j["clocks"]["sm_mhz"], j["config"]["last_counts_row"])
except Exception as ex:
    print("$name FAILED", ex)
PY
done
cat gpurun_out/exp.txt
