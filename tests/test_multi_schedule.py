"""Host logic of the multi-region hour loop (epirust_b200.multi.MultiRegion.run) on the CPU, with a recording stand-in for the
region engine: which hours are queued together, where the host waits, and that no hour runs before a host decision that
precedes it (Epidemiology::run_multi_engine, engine/src/epidemiology_simulation.rs:276-547; exchange hours
engine/src/transport/mpi_transport.rs:60-76; decision hours interventions/lockdown.rs:55,69-73, hospital.rs:55,70,
vaccination.rs:52)."""
import numpy as np
import pytest
import torch

from epirust_b200 import _ffi, multi


class FakeEngine:
    """Records the call sequence; Counts rows carry the hour so the placement of rows can be checked."""

    def __init__(self, vaccinate_at=(), unlock_at=None):
        self.calls, self.queued, self.vaccinate_at, self.unlock_at = [], [], tuple(vaccinate_at), unlock_at
        self.stream_ptr = 0

    def next_decision_hour(self, hour):
        d = (hour + 23) // 24 * 24
        for v in self.vaccinate_at:
            if v >= hour:
                d = min(d, v)
        if self.unlock_at is not None and self.unlock_at >= hour:
            d = min(d, self.unlock_at)
        return d

    def enqueue_hours(self, first, n):
        assert not self.queued or self.queued[-1] == first - 1, "queued hours must be consecutive"
        self.calls.append(("hours", first, n))
        self.queued += list(range(first, first + n))

    def enqueue_hour(self, hour):
        assert not self.queued or self.queued[-1] == hour - 1
        self.calls.append(("exchange_hour", hour))
        self.queued.append(-hour)  # exchange row: comes from finish_hour

    def travel_pack(self, hour, kind, ptr, stride, want_counts=True):
        self.calls.append(("pack", hour, kind))
        return np.zeros(2, np.uint32)

    def travel_unpack(self, hour, kind, ptr, stride, want_counts=True):
        self.calls.append(("unpack", hour, kind))

    def collect_hours(self):
        rows = np.array([[h, 0, 0, 0, 0, 0, 0] for h in self.queued if h > 0], np.uint32).reshape(-1, 7)
        self.calls.append(("collect", [h for h in self.queued if h > 0]))
        self.queued = []
        return rows

    def finish_hour(self, hour):
        self.calls.append(("finish", hour))
        return np.array([hour, 0, 0, 0, 0, 0, 0], np.uint32)

    def sync(self):
        pass


def make(monkeypatch, engines, migration=True, commute=True, start=48, end=336):
    plan = dict(n_regions=2, migration=np.ones((2, 2), np.uint32) if migration else None, commute=np.ones((2, 2), np.uint32) if commute else None,
                start_migration_hour=start, end_migration_hour=end)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: torch.empty(1))  # no device buffers on the CPU
    m = multi.MultiRegion.__new__(multi.MultiRegion)
    m.engines, m.on_outgoing, m.plan, m.kinds, m.exchange, m.R, m.stride = engines, None, plan, multi.exchange_hours(plan), None, 2, 8

    class Buf:
        def data_ptr(self):
            return 0

    class Send(list):
        def transpose(self, *_):
            return self

        def contiguous(self):
            return self

    m.send = Send([Buf() for _ in engines])
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: type("S", (), {"synchronize": lambda self: None, "cuda_stream": 0})())
    return m


def test_a_day_has_three_waits_and_rows_land_in_order(monkeypatch):
    e = FakeEngine()
    m = make(monkeypatch, [e])
    rows = m.run(49, 24)  # hours 49..72: day 3, inside the migration window
    assert rows[0, :, 0].tolist() == list(range(49, 73))
    assert [c for c in e.calls if c[0] in ("hours", "exchange_hour", "collect", "finish")] == [
        ("hours", 49, 6), ("exchange_hour", 55), ("collect", list(range(49, 55))), ("finish", 55),      # 07:00 commuters leave
        ("hours", 56, 9), ("exchange_hour", 65), ("collect", list(range(56, 65))), ("finish", 65),      # 17:00 commuters return
        ("hours", 66, 6), ("exchange_hour", 72), ("collect", list(range(66, 72))), ("finish", 72),      # 00:00 migrators
    ]
    kinds = [(c[1], c[2]) for c in e.calls if c[0] == "pack"]
    assert kinds == [(55, _ffi.TRAVEL_COMMUTE), (65, _ffi.TRAVEL_COMMUTE), (72, _ffi.TRAVEL_MIGRATE)]


def test_outside_the_migration_window_midnight_is_a_plain_decision_hour(monkeypatch):
    e = FakeEngine()
    m = make(monkeypatch, [e])
    m.run(18, 14)  # hours 18..31: midnight (24) is before start_migration_hour = 48
    seq = [c for c in e.calls if c[0] in ("hours", "exchange_hour", "collect", "finish")]
    assert seq == [("hours", 18, 7), ("collect", list(range(18, 25))),  # the host sees hour 24 before hour 25 runs (lockdown.rs:55)
                   ("hours", 25, 6), ("exchange_hour", 31), ("collect", list(range(25, 31))), ("finish", 31)]


@pytest.mark.parametrize("vaccinate_at,unlock_at", [((30,), None), ((), 54), ((54,), 60)])
def test_no_hour_runs_before_a_decision_that_precedes_it(monkeypatch, vaccinate_at, unlock_at):
    e = FakeEngine(vaccinate_at, unlock_at)
    m = make(monkeypatch, [e])
    rows = m.run(25, 48)
    assert rows[0, :, 0].tolist() == list(range(25, 73))
    decisions = set(vaccinate_at) | ({unlock_at} if unlock_at else set()) | {48, 72}
    seen = set()  # hours whose Counts the host has seen
    for c in e.calls:
        if c[0] in ("hours", "exchange_hour"):
            first = c[1]
            assert all(d in seen for d in decisions if d < first), f"{c} queued before a decision hour was collected"
        elif c[0] == "collect":
            seen |= set(c[1])
        elif c[0] == "finish":
            seen.add(c[1])


def test_decision_hour_right_before_an_exchange_is_collected_first(monkeypatch):
    e = FakeEngine(vaccinate_at=(54,))  # vaccination at 06:00, commuters leave at 07:00 (hour 55)
    m = make(monkeypatch, [e])
    m.run(49, 8)
    seq = [c for c in e.calls if c[0] in ("hours", "exchange_hour", "collect", "finish")]
    assert seq == [("hours", 49, 6), ("collect", list(range(49, 55))), ("exchange_hour", 55), ("collect", []), ("finish", 55), ("hours", 56, 1), ("collect", [56])]
