#!/bin/bash
# DRAM bytes of k_hour / k_commit at one work hour for each variant (ncu, a few metrics only)
mkdir -p gpurun_out
for v in ${VARIANTS:-0}; do
  EPI_LIB=$PWD/exp/lib_$v.so timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_hour|k_commit" -s ${SKIP:-80} -c ${COUNT:-4} --csv --log-file gpurun_out/expncu_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/expncu_$v.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/expncu_$v.csv", errors="ignore")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mn=h.index("Metric Name"); mv=h.index("Metric Value"); idc=h.index("ID")
cur={}
for r in rows[1:]:
    cur.setdefault((r[idc], r[kn].split("(")[0]), {})[r[mn]]=r[mv]
for k,m in cur.items(): print("variant $v", k, m)
PY
done
