"""Multi-region runs over the C ABI: one region engine per GPU, lock step, daily traveller exchange.

The hour loop (Epidemiology::run_multi_engine, engine/src/epidemiology_simulation.rs:276-547) and the Transport
(engine/src/transport/*.rs -> NCCL all-to-allv of packed records, or device copies when every region lives in one process) are
C++ behind include/epi.h (csrc/multi.cpp): epi_comm_init / epi_comm_init_local, epi_exchange, epi_run_multi_hours.  This
module is the ctypes harness the tests and the bench use; there is no Python on the product path (`engine-app -m mpi`).
"""
import ctypes as C

from . import _ffi
from .engine import EpiError, run_multi_hours


class MultiRegion:
    """The region engines this process hosts, sharing one communicator.

    engines: every region of the plan, in region order (local transport), or one engine together with (n_ranks, rank,
    unique_id) for an NCCL communicator (one process per GPU)."""

    def __init__(self, engines, n_ranks=None, rank=None, unique_id=None):
        self.engines = list(engines)
        L = _ffi.load()
        if unique_id is None:
            arr = (C.c_void_p * len(self.engines))(*[e.h for e in self.engines])
            if L.epi_comm_init_local(arr, len(self.engines)):
                raise EpiError(L.epi_last_error(self.engines[0].h).decode())
        else:
            (e,) = self.engines
            e.comm_init(n_ranks, rank, unique_id)

    def run(self, first_hour, n_hours, terminate_when_clear=False):
        """Hours first_hour .. first_hour + n_hours - 1 of every local region: rows[n_local, n_rows, 7]."""
        return run_multi_hours(self.engines, first_hour, n_hours, terminate_when_clear)

    def close(self):
        for e in self.engines:
            if getattr(e, "h", None):
                e.comm_destroy()


# ---- rendezvous of a torchrun-launched job (bench.py, tests): control plane only ------------------------------------------
def share_unique_id(dist, device="cpu"):
    """Rank 0 creates the NCCL unique id (epi_comm_unique_id) and broadcasts its 128 bytes over the job's torch.distributed
    group -- the role of the file that `engine-app -m mpi` passes its ranks (csrc/engine_app_main.cpp).  The traveller exchange
    itself never goes through torch.distributed: it is ncclSend/ncclRecv inside epi_exchange."""
    import torch

    from .engine import comm_unique_id

    t = torch.zeros(_ffi.COMM_ID_BYTES, dtype=torch.uint8)
    if dist.get_rank() == 0:
        t = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).clone()
    t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def max_over_ranks(dist, values, device="cpu"):
    """element-wise maximum of a list of floats over all ranks (device time of a multi-GPU measurement)"""
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]
