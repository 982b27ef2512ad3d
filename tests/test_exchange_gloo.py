"""The N > 1 plumbing on CPU, world_size-2 gloo: how the ranks of a torchrun-launched job (bench.py) meet -- rank 0's NCCL
unique id reaches every rank (epirust_b200.multi.share_unique_id), timings are reduced as the maximum over ranks, every rank
derives the same exchange schedule from the plan (epi_multi_schedule_trace: the real C++ hour loop with recording stand-ins)
and, asked to join a communicator without a GPU, fails loudly with EPI_ERR_* instead of hanging or falling back."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    from epirust_b200 import _ffi
    from epirust_b200.engine import multi_schedule_trace
    from epirust_b200.multi import max_over_ranks, share_unique_id

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = share_unique_id(dist)
        ids = [None] * world
        dist.all_gather_object(ids, uid)
        ok = len(uid) == _ffi.COMM_ID_BYTES and any(uid) and all(i == uid for i in ids)
        ms = max_over_ranks(dist, [10.0 + rank, 5.0 - rank])
        ok = ok and ms == [10.0 + world - 1, 5.0]
        plan = dict(n_regions=world, migration=np.ones((world, world), np.uint32), commute=np.ones((world, world), np.uint32), start_migration_hour=48, end_migration_hour=336)
        calls = multi_schedule_trace(plan, 1, 96)
        mine = [c for c in calls if c[0] == "exchange"]
        every = [None] * world
        dist.all_gather_object(every, mine)
        ok = ok and len(mine) == 2 * 4 + 2 and all(x == mine for x in every)  # 07:00 + 17:00 every day, midnight of days 3 and 4 (hours 72, 96)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_rendezvous_and_schedule_world_size_2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
