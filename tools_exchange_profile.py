"""Per-phase wall clock of the exchange hours under torchrun (one region per GPU):
    python -m torch.distributed.run --nproc-per-node N tools_exchange_profile.py [workload] [days]"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from epirust_b200.engine import Engine, make_config
from epirust_b200.multi import MultiRegion, DistExchange
from bench import WORKLOADS, travel_plan_for

wl = sys.argv[1] if len(sys.argv) > 1 else '10m'
days = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
kw = dict(WORKLOADS[wl]); n = kw['n_agents']
plan = travel_plan_for(world, n)
eng = Engine(make_config(hours=4000, **kw), seed=1 + rank, device=local, region=rank, plan=plan, extra_capacity=n // 25)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
T = {}
def timed(name, f):
    t = time.perf_counter(); r = f(); eng.sync(); torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t; return r
with torch.cuda.stream(stream):
    m = MultiRegion([eng], plan, exchange=DistExchange(torch.device('cuda', local)), stride_records=2 * (n // 1000) + 4096)
    m.run(1, 72)
    hour = 73
    torch.cuda.synchronize(); dist.barrier(); t_all = time.perf_counter()
    for day in range(days):
        for h in range(24):
            x = hour + h
            from epirust_b200.multi import exchange_kind
            kind = exchange_kind(plan, m.kinds, x)
            if kind is not None:
                tag = 'mig' if kind == 0 else 'com'
                timed('step', lambda: eng.enqueue_hour(x))
                counts = timed('pack_' + tag, lambda: eng.travel_pack(x, kind, m.send[0].data_ptr(), m.stride))
                recv = timed('nccl_' + tag, lambda: m.exchange.exchange(m.send[0]))
                timed('unpack_' + tag, lambda: eng.travel_unpack(x, kind, recv.data_ptr(), m.stride))
                timed('finish', lambda: eng.finish_hour(x))
            else:
                timed('plain_hours', lambda: eng.simulate_hours(x, 1))
        hour += 24
    torch.cuda.synchronize(); dist.barrier(); t_all = time.perf_counter() - t_all
if rank == 0:
    print(wl, 'world', world, {k: round(v * 1e3 / days, 3) for k, v in T.items()}, 'ms per simulated day; day total %.3f ms' % (t_all * 1e3 / days), 'travellers/exchange', int(counts.sum()))
dist.destroy_process_group()
