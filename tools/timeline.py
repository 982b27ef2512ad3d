"""GPU timeline of a multi-region day (EPI_TRACE=1): when each hour's commit pass starts and the stamps of the exchange kernels.
    EPI_TRACE=1 python -m torch.distributed.run --nproc-per-node N tools/timeline.py [workload] [days]"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epirust_b200.engine import Engine, make_config
from epirust_b200.multi import MultiRegion, share_unique_id
from bench import WORKLOADS, travel_plan_for

wl = sys.argv[1] if len(sys.argv) > 1 else "10m"
days = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
kw = dict(WORKLOADS[wl]); n = kw["n_agents"]
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    plan = travel_plan_for(world, n)
    eng = Engine(make_config(hours=2000, **kw), seed=1 + rank, device=local, region=rank, plan=plan, extra_capacity=max(32768, 2 * (world - 1) * (n // 1000 + n // 2000)))
    m = MultiRegion([eng], n_ranks=world, rank=rank, unique_id=share_unique_id(dist, device=torch.device("cuda", local)))
    run = lambda h, k: m.run(h, k)
else:
    eng = Engine(make_config(hours=2000, **kw), seed=1)
    run = lambda h, k: eng.simulate_hours(h, k)
W = 3
for d in range(W):
    run(24 * d + 1, 24)
eng.debug_trace()
torch.cuda.synchronize()
if world > 1: dist.barrier()
for d in range(W, W + days):
    run(24 * d + 1, 24)
tr = eng.debug_trace()
names = {1: "commit", 2: "leave>", 3: "leave<", 4: "arrive>", 5: "arrived", 6: "arrive<"}
if rank in (0, world - 1):
    t0 = tr[0][1]
    prev = t0
    out = []
    for tag, t, hour in tr:
        out.append("r%d %-8s h=%4d (%2d)  t=%9.1f us  +%7.1f" % (rank, names.get(tag, tag), hour, hour % 24, (t - t0) / 1e3, (t - prev) / 1e3))
        prev = t
    open("gpurun_out/timeline_r%d.txt" % rank, "w").write("\n".join(out) + "\n")
    print("\n".join(out[: 70]))
if world > 1: dist.destroy_process_group()
