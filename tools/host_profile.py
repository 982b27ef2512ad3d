import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from epirust_b200.engine import Engine, make_config
from epirust_b200.multi import MultiRegion, DistExchange
from bench import WORKLOADS, travel_plan_for
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dist.init_process_group('nccl', device_id=torch.device('cuda', local))
kw = dict(WORKLOADS['10m']); n = kw['n_agents']; plan = travel_plan_for(world, n)
eng = Engine(make_config(hours=8000, **kw), seed=1 + rank, device=local, region=rank, plan=plan, extra_capacity=max(32768, 2 * (world - 1) * (n // 1000 + n // 2000)))
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
T = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); T[name] = T.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(obj, name, g)
with torch.cuda.stream(stream):
    m = MultiRegion([eng], plan, exchange=DistExchange(torch.device('cuda', local)), stride_records=2 * (n // 1000) + 4096)
    m.run(1, 72)
    for nme in ('enqueue_hours', 'enqueue_hour', 'travel_pack', 'travel_unpack', 'collect_hours', 'finish_hour', 'next_decision_hour'): wrap(eng, nme)
    wrap(m.exchange, 'exchange')
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    days = 6
    m.run(73, 24 * days)
    torch.cuda.synchronize(); total = time.perf_counter() - t0
if rank == 0:
    print('host ms/day:', {k: round(v * 1e3 / days, 3) for k, v in T.items()}, 'total', round(total * 1e3 / days, 3))
    print('launches', eng.launch_count())
dist.destroy_process_group()
