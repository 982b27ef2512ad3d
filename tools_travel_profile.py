"""Where does an exchange hour spend its time?  2 regions on one GPU, wall clock per phase."""
import sys, time
import numpy as np
import torch
from epirust_b200.engine import Engine, make_config
from epirust_b200.multi import MultiRegion
from epirust_b200 import _ffi
sys.path.insert(0, '.')
from bench import WORKLOADS, travel_plan_for

wl = sys.argv[1] if len(sys.argv) > 1 else '2m'
kw = dict(WORKLOADS[wl]); n = kw['n_agents']; R = 2
plan = travel_plan_for(R, n)
cfg = make_config(hours=2000, **kw)
engines = [Engine(cfg, seed=1 + r, device=0, region=r, plan=plan, extra_capacity=n // 25) for r in range(R)]
m = MultiRegion(engines, plan, max_records=1 << 18)
m.run(1, 72)
def sync():
    for e in engines: e.sync()
    torch.cuda.synchronize()
T = {}
def timed(name, f):
    sync(); t = time.perf_counter(); r = f(); sync(); T[name] = T.get(name, 0) + time.perf_counter() - t; return r
hour = 73
for day in range(3):
    for h in range(24):
        x = hour + h
        if x % 24 in m.kinds:
            kind = m.kinds[x % 24]
            timed('step', lambda: [e.step(x) for e in engines])
            outs = timed('pack_%d' % kind, lambda: [e.travel_pack(x, kind, b.data_ptr(), m.max_records) for e, b in zip(engines, m.send)])
            from epirust_b200.multi import split_records
            parts = [split_records(b, c) for b, c in zip(m.send, outs)]
            for r, e in enumerate(engines):
                ci = np.array([outs[s][r] for s in range(R)], np.uint32)
                recv = timed('cat', lambda: torch.cat([parts[s][r] for s in range(R)], dim=0).contiguous())
                timed('unpack_%d' % kind, lambda: e.travel_unpack(x, kind, recv.data_ptr(), ci))
            timed('finish', lambda: [e.finish_hour(x) for e in engines])
        else:
            timed('plain_hours', lambda: [e.simulate_hours(x, 1) for e in engines])
    hour += 24
print({k: round(v * 1e3 / 3, 3) for k, v in T.items()}, 'ms per simulated day, both regions')
