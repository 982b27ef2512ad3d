#!/bin/bash
# kernel-level view of the exchange: ncu launch list of a 2-region run on ONE GPU (LocalExchange), travel kernels only
mkdir -p gpurun_out
cat > /tmp/x.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from epirust_b200.engine import Engine, make_config
from epirust_b200.multi import MultiRegion
from bench import WORKLOADS, travel_plan_for
kw = dict(WORKLOADS['10m']); n = kw['n_agents']; R = 2
plan = travel_plan_for(R, n)
engines = [Engine(make_config(hours=2000, **kw), seed=1 + r, device=0, region=r, plan=plan, extra_capacity=n // 25) for r in range(R)]
m = MultiRegion(engines)
m.run(1, 73)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_travel|k_occ" --csv --log-file gpurun_out/travel_kernels.csv python /tmp/x.py > gpurun_out/travel_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/travel_kernels.csv', errors='ignore')) if len(r)>10]
h=rows[0]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[kn].split('(')[0].replace('void ','').replace('epi::','')
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[mv].replace(',',''))/1e3
tot=sum(v[1] for v in agg.values())
for k,v in agg.items(): print('%-40s n=%4d total %8.1f us  avg %7.1f us' % (k, v[0], v[1], v[1]/v[0]))
print('total', tot, 'us over both regions, 3 simulated days')
PY
