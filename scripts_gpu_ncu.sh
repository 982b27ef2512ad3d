#!/bin/bash
# full ncu capture: k_hour at h=11 (work hour), h=15, h=19 (home hour); ARGS: skip counts
mkdir -p gpurun_out
for s in ${SKIPS:-40 48}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_hour} -s $s -c 1 -f -o gpurun_out/prof_${KERNEL:-k_hour}_$s python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$s.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
