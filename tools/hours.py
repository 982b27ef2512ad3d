"""Per hour of day: average duration of the hour's agent kernels and of its commit pass (CUDA events around every launch, graphs
off) over D simulated days after W warm-up days.  python tools/hours.py [workload] [days] [warmup]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from epirust_b200.engine import Engine, make_config

wl = sys.argv[1] if len(sys.argv) > 1 else "10m"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kw = dict(bench.WORKLOADS[wl])
eng = Engine(make_config(hours=24 * (D + W) + 1, **kw), seed=1)
eng.simulate_hours(1, 24 * W)
eng.set_kernel_timing(True)
eng.simulate_hours(24 * W + 1, 24 * D)
ht = eng.hour_times()
tot = 0.0
for h in sorted(ht, key=lambda h: (h - 7) % 24):
    a, na, c, nc = ht[h]
    tot += (a + c) / D
    print("h=%2d hour %6.1f us  commit %6.1f us  pass %6.1f us" % (h, 1e3 * a / max(na, 1), 1e3 * c / max(nc, 1), 1e3 * (a / max(na, 1) + c / max(nc, 1))))
kt = eng.kernel_times()
print("active passes per day %.3f ms; sleep %.1f us; tiles %s  tile_hours %d" % (tot, 1e3 * kt["sleep"][0] / max(1, kt["sleep"][1]), os.environ.get("EPI_TILES", "1"), eng.tile_hours))
