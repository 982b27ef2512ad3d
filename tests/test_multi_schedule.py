"""Host logic of the multi-region hour loop (epi_run_multi_hours, csrc/multi.cpp) on the CPU: the real C++ scheduling code
driven with recording stand-ins for the region engine and the transport (epi_multi_schedule_trace) -- which hours are queued
together, where the host waits, and that no hour runs before a host decision that precedes it
(Epidemiology::run_multi_engine, engine/src/epidemiology_simulation.rs:276-547; exchange hours
engine/src/transport/mpi_transport.rs:60-76; decision hours interventions/lockdown.rs:55,69-73, hospital.rs:55,70,
vaccination.rs:52), and the orchestrator's termination rule (orchestrator/src/ticks.rs:175-180, KAT :208-229)."""
import json
import os

import numpy as np
import pytest

from epirust_b200 import _ffi
from epirust_b200.engine import multi_schedule_trace, should_terminate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def plan(migration=True, commute=True, start=48, end=336):
    return dict(n_regions=2, migration=np.ones((2, 2), np.uint32) if migration else None, commute=np.ones((2, 2), np.uint32) if commute else None,
                start_migration_hour=start, end_migration_hour=end)


def test_a_day_has_one_wait_and_rows_land_in_order():
    calls = multi_schedule_trace(plan(), 49, 24)  # hours 49..72: day 3, inside the migration window
    assert [c for c in calls if c[0] != "exchange"] == [
        ("hours", 49, 6), ("exchange_hour", 55),      # 07:00 commuters leave
        ("hours", 56, 9), ("exchange_hour", 65),      # 17:00 commuters return
        ("hours", 66, 6), ("exchange_hour", 72),      # 00:00 migrators
        ("collect", list(range(49, 73))),             # the only wait of the day: midnight is a decision hour (lockdown.rs:55, hospital.rs:55,70)
    ]
    assert [c for c in calls if c[0] == "exchange"] == [("exchange", 55, _ffi.TRAVEL_COMMUTE), ("exchange", 65, _ffi.TRAVEL_COMMUTE), ("exchange", 72, _ffi.TRAVEL_MIGRATE)]
    # the exchange is queued right behind its hour's kernels, and the next hours right behind the exchange: the host does not wait
    assert calls.index(("exchange", 55, _ffi.TRAVEL_COMMUTE)) == calls.index(("exchange_hour", 55)) + 1
    assert calls.index(("hours", 56, 9)) == calls.index(("exchange", 55, _ffi.TRAVEL_COMMUTE)) + 1


def test_outside_the_migration_window_midnight_is_a_plain_decision_hour():
    calls = multi_schedule_trace(plan(), 18, 14)  # hours 18..31: midnight (24) is before start_migration_hour = 48
    assert [c for c in calls if c[0] != "exchange"] == [
        ("hours", 18, 7), ("collect", list(range(18, 25))),  # the host sees hour 24 before hour 25 runs (lockdown.rs:55)
        ("hours", 25, 6), ("exchange_hour", 31), ("collect", list(range(25, 32)))]


@pytest.mark.parametrize("vaccinate_at,unlock_at", [((30,), 0), ((), 54), ((54,), 60)])
def test_no_hour_runs_before_a_decision_that_precedes_it(vaccinate_at, unlock_at):
    calls = multi_schedule_trace(plan(), 25, 48, vaccinate_hours=vaccinate_at, unlock_hour=unlock_at)
    decisions = set(vaccinate_at) | ({unlock_at} if unlock_at else set()) | {48, 72}
    seen, rows = set(), []  # hours whose Counts the host has seen
    for c in calls:
        if c[0] in ("hours", "exchange_hour"):
            first = c[1]
            assert all(d in seen for d in decisions if d < first), f"{c} queued before a decision hour was collected"
        elif c[0] == "collect":
            seen |= set(c[1])
            rows += c[1]
    assert rows == list(range(25, 73))


def test_decision_hour_right_before_an_exchange_is_collected_first():
    calls = multi_schedule_trace(plan(), 49, 8, vaccinate_hours=(54,))  # vaccination at 06:00, commuters leave at 07:00 (hour 55)
    assert [c for c in calls if c[0] != "exchange"] == [
        ("hours", 49, 6), ("collect", list(range(49, 55))), ("exchange_hour", 55), ("hours", 56, 1), ("collect", [55, 56])]


def test_a_decision_at_an_exchange_hour_waits_after_the_exchange():
    calls = multi_schedule_trace(plan(), 49, 10, vaccinate_hours=(55,))  # vaccination at 07:00 = the hour the commuters leave
    assert calls == [("hours", 49, 6), ("exchange_hour", 55), ("exchange", 55, _ffi.TRAVEL_COMMUTE), ("collect", list(range(49, 56))),
                     ("hours", 56, 3), ("collect", [56, 57, 58])]


def test_every_rank_issues_the_same_collectives_whatever_its_own_decision_hours():
    """Lock step without a barrier: the sequence of exchanges depends on the travel plan and the hour only, never on a region's own
    interventions -- otherwise two ranks would wait for each other in different collectives."""
    want = None
    for vac, unlock in (((), 0), ((30, 55, 100), 0), ((7, 17, 24), 65), ((54,), 31)):
        calls = multi_schedule_trace(plan(start=24, end=200), 1, 240, vaccinate_hours=vac, unlock_hour=unlock)
        x = [c for c in calls if c[0] == "exchange"]
        want = want or x
        assert x == want and len(x) == 10 * 2 + 7  # 07:00 and 17:00 of ten days, midnight of hours 48..192


def test_termination_rule_known_answers():
    """orchestrator/src/ticks.rs:208-229 (should_terminate_when_exposed_and_infected_and_hospitalized_are_zero)"""
    kat = json.load(open(os.path.join(GOLDEN, "reference_kats.json")))["orchestrator_should_terminate"]
    for case in kat["cases"]:
        assert should_terminate(np.array(case["acks"], np.uint32)) is case["terminate"], case
