"""The N > 1 plumbing on CPU: world_size-2 gloo run of the padded all-to-all that carries the traveller records
(epirust_b200.multi.DistExchange), with CPU tensors standing in for the device buffers."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    from epirust_b200.multi import DistExchange, REC_WORDS

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = DistExchange(torch.device("cpu"))
        stride = 16
        # rank r sends (r + 1) * (d + 2) records to rank d; header word 0 = count; record word 0 = sender, 1 = destination, 2 = index
        counts = np.array([(rank + 1) * (d + 2) if d != rank else 0 for d in range(world)], np.int64)
        send = torch.zeros((world, stride, REC_WORDS), dtype=torch.int32)
        for d in range(world):
            send[d, 0, 0] = int(counts[d])
            for k in range(int(counts[d])):
                send[d, 1 + k, :3] = torch.tensor([rank, d, k], dtype=torch.int32)
        recv = x.exchange(send)
        want_in = np.array([(s + 1) * (rank + 2) if s != rank else 0 for s in range(world)], np.int64)
        ok = recv.shape == send.shape and recv[:, 0, 0].tolist() == want_in.tolist()
        for s in range(world):
            part = recv[s, 1:1 + int(want_in[s])]
            ok = ok and bool((part[:, 0] == s).all()) and bool((part[:, 1] == rank).all()) and part[:, 2].tolist() == list(range(int(want_in[s])))
        total = x.all_reduce_sum([int(counts.sum())])[0]
        ok = ok and total == sum((r + 1) * (d + 2) for r in range(world) for d in range(world) if d != r)
        # an exchange in which nobody travels still works (zero headers)
        recv2 = x.exchange(torch.zeros((world, stride, REC_WORDS), dtype=torch.int32))
        ok = ok and int(recv2[:, 0, 0].sum()) == 0
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_all_to_allv_of_traveller_records_world_size_2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
