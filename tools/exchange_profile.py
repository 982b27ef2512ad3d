"""Where a multi-region day goes (one region per GPU, torchrun): wall clock per simulated day of
  full      MultiRegion.run as benchmarked
  no_nccl   the same with the collective replaced by a device copy of the send buffer (every region receives its own leavers' layout)
  no_travel the same cut points (segments, collect, finish) without pack / collective / unpack
  one_graph plain 24-hour days through epi_simulate_hours (the single-region path)
    python -m torch.distributed.run --nproc-per-node N tools/exchange_profile.py [workload] [days]"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epirust_b200.engine import Engine, make_config
from epirust_b200.multi import MultiRegion, DistExchange, exchange_kind
from bench import WORKLOADS, travel_plan_for

wl = sys.argv[1] if len(sys.argv) > 1 else '10m'
days = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
kw = dict(WORKLOADS[wl]); n = kw['n_agents']
plan = travel_plan_for(world, n)
eng = Engine(make_config(hours=8000, **kw), seed=1 + rank, device=local, region=rank, plan=plan, extra_capacity=max(32768, 2 * (world - 1) * (n // 1000 + n // 2000)))
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
out = {}
def timed(name, f, hour):
    eng.reset(); m.run(1, 72)  # every variant times the same simulated days
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter(); f(hour); torch.cuda.synchronize(); dist.barrier()
    out[name] = (time.perf_counter() - t) * 1e3 / days
with torch.cuda.stream(stream):
    m = MultiRegion([eng], plan, exchange=DistExchange(torch.device('cuda', local)), stride_records=2 * (n // 1000) + 4096)
    m.run(1, 72)
    hour = 73
    timed('full', lambda h: m.run(h, 24 * days), hour)
    class Local:
        def exchange(self, send):
            return send.clone()
    real = m.exchange
    m.exchange = Local()
    timed('no_nccl', lambda h: m.run(h, 24 * days), hour)
    m.exchange = real
    def no_travel(h0):
        h, last = h0, h0 + 24 * days - 1
        while h <= last:
            x = m.next_exchange_hour(h, last)
            seg_end = min(last, eng.next_decision_hour(h), (x - 1) if x is not None else last)
            if seg_end >= h: eng.enqueue_hours(h, seg_end - h + 1)
            ex = x is not None and x == seg_end + 1 and not (seg_end >= h and seg_end == eng.next_decision_hour(h))
            if ex: eng.enqueue_hour(x)
            eng.collect_hours()
            if ex: eng.finish_hour(x)
            h = (x if ex else seg_end) + 1
    timed('no_travel', no_travel, hour)
    timed('one_graph', lambda h: eng.simulate_hours(h, 24 * days), hour)
if rank == 0:
    print(wl, 'world', world, {k: round(v, 3) for k, v in out.items()}, 'ms per simulated day')
dist.destroy_process_group()
