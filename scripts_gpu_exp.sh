#!/bin/bash
# A/B kernel experiments: bench every exp/lib_<n>.so named in $VARIANTS on the 10M workload (and 1M when WL1M=1)
mkdir -p gpurun_out
: > gpurun_out/exp.txt
for v in ${VARIANTS:-0 1}; do
  for wl in ${WLS:-10m}; do
    EPI_LIB=$PWD/exp/lib_$v.so timeout 300 python bench.py --workload $wl --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/exp_${v}_$wl.json 2> gpurun_out/exp_${v}_$wl.err
    python - <<PY >> gpurun_out/exp.txt
import json
try:
    j = json.loads(open("gpurun_out/exp_${v}_$wl.json").read().strip().splitlines()[-1])
    print("variant $v $wl value %.4e ms/day %.3f" % (j["value"], j["ms_per_step"]), {k: round(x, 4) for k, x in j["roofline"]["per_kernel_ms"].items()}, j["config"]["last_counts_row"])
except Exception as ex:
    print("variant $v $wl FAILED", ex)
PY
  done
done
cat gpurun_out/exp.txt
