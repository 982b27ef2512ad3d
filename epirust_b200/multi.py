"""Multi-region runs: one region engine per GPU, lock step, daily traveller exchange.

Mirror of Epidemiology::run_multi_engine (engine/src/epidemiology_simulation.rs:276-547) with the MPI / Kafka transport
(engine/src/transport/*.rs) replaced by an all-to-allv of packed traveller records between the GPUs:

  * DistExchange  -- one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch; gloo for CPU tests of the
                     plumbing): ONE all_to_all_single with equal splits over padded segments whose headers carry the counts.
  * exchange=None -- every region engine lives in this process (tests; several regions on one GPU): a device transpose.

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


def exchange_kind(plan, kinds, hour):
    """Kind of the exchange at `hour`, or None.  A migration hour outside the migration window moves nobody in any region
    (Citizen::can_migrate, citizen/mod.rs:460-462), so every rank skips it alike."""
    kind = kinds.get(hour % 24)
    if kind == _ffi.TRAVEL_MIGRATE and not (int(plan.get("start_migration_hour", 0)) < hour < int(plan.get("end_migration_hour", 0))):
        return None
    return kind


class DistExchange:
    """all-to-all over torch.distributed; `device` is the torch device of the record buffers.  The buffers are padded (one
    fixed-size segment per peer, the record count in the segment header), so ONE collective with equal splits moves counts
    and payload together and no count exchange / host round trip precedes it."""

    def __init__(self, device, group=None):
        import torch.distributed as dist

        self.dist, self.group, self.device = dist, group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._recv = None

    def exchange(self, send_buf):
        """send_buf: [world, stride, REC_WORDS] int32, segment d addressed to rank d.  Returns recv of the same shape: segment s = what rank s sent."""
        if self._recv is None or self._recv.shape != send_buf.shape:
            self._recv = torch.empty_like(send_buf)
        self.dist.all_to_all_single(self._recv.view(-1), send_buf.view(-1), group=self.group)
        return self._recv

    def all_reduce_sum(self, values):
        t = torch.as_tensor(np.asarray(values, np.int64), device=self.device)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()


class MultiRegion:
    """R region engines hosted by this process (R == 1 per process under torchrun)."""

    def __init__(self, engines, plan, exchange=None, stride_records=1 << 15, on_outgoing=None):
        """stride_records: capacity of one (source, destination) segment, header included.
        on_outgoing(hour, kind, send_buf, counts): called per local engine after its leavers were packed (send_buf: [n_regions, stride, 8]
        int32 device tensor, segment d = header + records for region d; counts: numpy[n_regions]) -- the hook of
        Listener::outgoing_migrators_added."""
        self.engines = engines
        self.on_outgoing = on_outgoing
        self.plan = plan
        self.kinds = exchange_hours(plan)
        self.exchange = exchange  # None: all regions are local
        self.R = int(plan["n_regions"])
        self.stride = int(stride_records)
        dev = torch.device("cuda", torch.cuda.current_device())
        # [local engine, destination region, record, word]
        self.send = torch.zeros((len(engines), self.R, self.stride, REC_WORDS), dtype=torch.int32, device=dev)

    def next_exchange_hour(self, hour, last_hour):
        for h in range(hour, last_hour + 1):
            if exchange_kind(self.plan, self.kinds, h) is not None:
                return h
        return None

    def _do_exchange(self, hour, kind):
        # Without a listener nothing on the host needs the counts: pack, collective and unpack are queued back to back on the
        # stream and finish_hour settles the host side with one synchronisation.
        deferred = self.on_outgoing is None
        for i, e in enumerate(self.engines):
            counts = e.travel_pack(hour, kind, self.send[i].data_ptr(), self.stride, want_counts=not deferred)
            if self.on_outgoing is not None:
                self.on_outgoing(hour, kind, self.send[i], counts)
        if self.exchange is None:
            # every region is local: region r receives segment r of every source, in source order
            for e in self.engines:
                e.sync()
            recv = self.send.transpose(0, 1).contiguous()
            torch.cuda.current_stream().synchronize()
            for r, e in enumerate(self.engines):
                e.travel_unpack(hour, kind, recv[r].data_ptr(), self.stride, want_counts=not deferred)
        else:
            (e,) = self.engines
            if e.stream_ptr != torch.cuda.current_stream().cuda_stream:  # the collective runs on another stream than the engine's
                e.sync()
            recv = self.exchange.exchange(self.send[0])
            if e.stream_ptr != torch.cuda.current_stream().cuda_stream:
                torch.cuda.current_stream().synchronize()
            e.travel_unpack(hour, kind, recv.data_ptr(), self.stride, want_counts=not deferred)

    def run(self, first_hour, n_hours, rows_out=None):
        """Hours first_hour .. first_hour + n_hours - 1 of every local region.  Returns rows[n_local, n_hours, 7].

        The host waits for the device only where it has to look at Counts: after an exchange hour (epi_finish_hour) and at a
        decision hour of process_interventions (start of day, vaccination hour, unlock hour).  Everything in between -- the
        plain hours before an exchange, the exchange hour's kernels, pack, collective and unpack -- is queued back to back."""
        rows = rows_out if rows_out is not None else np.zeros((len(self.engines), n_hours, 7), np.uint32)
        hour, last = first_hour, first_hour + n_hours - 1
        while hour <= last:
            x = self.next_exchange_hour(hour, last)
            decision = min(e.next_decision_hour(hour) for e in self.engines)
            seg_end = min(last, decision, (x - 1) if x is not None else last)  # last plain hour queued in this round
            if seg_end >= hour:
                for e in self.engines:
                    e.enqueue_hours(hour, seg_end - hour + 1)
            exchange_now = x is not None and x == seg_end + 1 and not (seg_end >= hour and seg_end == decision)
            if exchange_now:
                for e in self.engines:
                    e.enqueue_hour(x)
                self._do_exchange(x, exchange_kind(self.plan, self.kinds, x))
            for i, e in enumerate(self.engines):
                got = e.collect_hours()
                if len(got):
                    rows[i, hour - first_hour:hour - first_hour + len(got)] = got
                if exchange_now:
                    rows[i, x - first_hour] = e.finish_hour(x)
            hour = (x if exchange_now else seg_end) + 1
        return rows

    def active_cases_everywhere(self, last_rows):
        """orchestrator/src/ticks.rs:175-180: the run may terminate when no region has exposed / infected / hospitalized agents."""
        local = np.array([int(r[2]) + int(r[3]) + int(r[4]) for r in last_rows]).sum()
        if self.exchange is not None:
            local = int(self.exchange.all_reduce_sum([local])[0])
        return local > 0
