"""Multi-region runs: one region engine per GPU, lock step, daily traveller exchange.

Mirror of Epidemiology::run_multi_engine (engine/src/epidemiology_simulation.rs:276-547) with the MPI / Kafka transport
(engine/src/transport/*.rs) replaced by an all-to-allv of packed traveller records between the GPUs:

  * DistExchange  -- one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch; gloo for CPU tests of the
                     plumbing): counts via all_to_all_single, then the records via all_to_all_single with split sizes.
  * LocalExchange -- every region engine lives in this process (tests; several regions on one GPU): device-to-device
                     copies.

The orchestrator's barrier (orchestrator/src/ticks.rs:35-89) is implicit in the collective; its global termination rule
(sum of exposed + infected + hospitalized over the regions == 0, ticks.rs:175-180) is `active_cases_everywhere`.
"""
import numpy as np
import torch

from . import _ffi

REC_WORDS = _ffi.TRAVEL_RECORD_BYTES // 4


def exchange_hours(plan):
    """Hours of the day with an exchange and their kind (mpi_transport.rs:60-76)."""
    kinds = {}
    if plan.get("migration") is not None:
        kinds[0] = _ffi.TRAVEL_MIGRATE
    if plan.get("commute") is not None:
        kinds[7] = _ffi.TRAVEL_COMMUTE
        kinds[17] = _ffi.TRAVEL_COMMUTE
    return kinds


def split_records(buf, counts):
    """Views of a flat record buffer per peer."""
    out, at = [], 0
    for c in counts:
        out.append(buf[at:at + int(c)])
        at += int(c)
    return out


class DistExchange:
    """all-to-allv over torch.distributed; `device` is the torch device of the record buffers."""

    def __init__(self, device, group=None):
        import torch.distributed as dist

        self.dist, self.group, self.device = dist, group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def exchange(self, send_buf, send_counts):
        """send_buf: [n, REC_WORDS] int32 tensor grouped by destination rank; send_counts: numpy[world].  Returns (recv_buf, recv_counts)."""
        d = self.dist
        sc = torch.as_tensor(np.asarray(send_counts, np.int64), device=self.device)
        rc = torch.empty_like(sc)
        d.all_to_all_single(rc, sc, group=self.group)
        recv_counts = rc.cpu().numpy()
        n_in = int(recv_counts.sum())
        recv = torch.empty((n_in, REC_WORDS), dtype=torch.int32, device=self.device)
        d.all_to_all_single(recv, send_buf[: int(np.sum(send_counts))], output_split_sizes=[int(c) for c in recv_counts],
                            input_split_sizes=[int(c) for c in send_counts], group=self.group)
        return recv, recv_counts.astype(np.uint32)

    def all_reduce_sum(self, values):
        t = torch.as_tensor(np.asarray(values, np.int64), device=self.device)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()


class MultiRegion:
    """R region engines hosted by this process (R == 1 per process under torchrun)."""

    def __init__(self, engines, plan, exchange=None, max_records=1 << 17, on_outgoing=None):
        """on_outgoing(hour, kind, send_buf, counts): called per local engine after its leavers were packed (send_buf: [n, 8] int32
        device tensor grouped by destination; counts: numpy[n_regions]) -- the hook of Listener::outgoing_migrators_added."""
        self.engines = engines
        self.on_outgoing = on_outgoing
        self.plan = plan
        self.kinds = exchange_hours(plan)
        self.exchange = exchange  # None: all regions are local
        self.R = int(plan["n_regions"])
        dev = torch.device("cuda", torch.cuda.current_device())
        self.send = [torch.zeros((max_records, REC_WORDS), dtype=torch.int32, device=dev) for _ in engines]
        self.max_records = max_records

    def next_exchange_hour(self, hour, last_hour):
        for h in range(hour, last_hour + 1):
            if h % 24 in self.kinds:
                return h
        return None

    def _do_exchange(self, hour, kind):
        outs = []
        for e, buf in zip(self.engines, self.send):
            outs.append(e.travel_pack(hour, kind, buf.data_ptr(), self.max_records))
            if self.on_outgoing is not None:
                e.sync()
                self.on_outgoing(hour, kind, buf, outs[-1])
        if self.exchange is None:
            # local all-to-allv: region r receives, in source order, what every source addressed to it
            parts = [split_records(buf, c) for buf, c in zip(self.send, outs)]
            for r, e in enumerate(self.engines):
                counts_in = np.array([outs[s][r] for s in range(self.R)], np.uint32)
                if counts_in.sum() == 0:
                    continue
                recv = torch.cat([parts[s][r] for s in range(self.R)], dim=0).contiguous()
                torch.cuda.current_stream().synchronize()
                e.travel_unpack(hour, kind, recv.data_ptr(), counts_in)
        else:
            (e,), (buf,), (counts,) = self.engines, self.send, outs
            recv, counts_in = self.exchange.exchange(buf, counts)
            torch.cuda.current_stream().synchronize()
            if counts_in.sum():
                e.travel_unpack(hour, kind, recv.data_ptr(), counts_in)

    def run(self, first_hour, n_hours, rows_out=None):
        """Hours first_hour .. first_hour + n_hours - 1 of every local region.  Returns rows[n_local, n_hours, 7]."""
        rows = rows_out if rows_out is not None else np.zeros((len(self.engines), n_hours, 7), np.uint32)
        hour, last = first_hour, first_hour + n_hours - 1
        while hour <= last:
            x = self.next_exchange_hour(hour, last)
            stop = x if x is not None else last + 1
            if stop > hour:
                for i, e in enumerate(self.engines):
                    got, _ = e.simulate_hours(hour, stop - hour, stop_rule=False, out=rows[i, hour - first_hour:stop - first_hour])
            if x is None:
                break
            for e in self.engines:
                e.step(x)
            for e in self.engines:
                e.sync()
            self._do_exchange(x, self.kinds[x % 24])
            for i, e in enumerate(self.engines):
                rows[i, x - first_hour] = e.finish_hour(x)
            hour = x + 1
        return rows

    def active_cases_everywhere(self, last_rows):
        """orchestrator/src/ticks.rs:175-180: the run may terminate when no region has exposed / infected / hospitalized agents."""
        local = np.array([int(r[2]) + int(r[3]) + int(r[4]) for r in last_rows]).sum()
        if self.exchange is not None:
            local = int(self.exchange.all_reduce_sum([local])[0])
        return local > 0
