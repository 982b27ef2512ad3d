"""GPU parity of multi-region runs: R region engines on one GPU exchanging travellers device-to-device, against the
multi-region oracle (same seeds, same deterministic conventions): every hour's Counts row of every region, and the full
agent state by slot, bit-exact."""
import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200.engine import Engine, make_config, STATE_FIELDS
from epirust_b200.multi import MultiRegion

pytestmark = pytest.mark.gpu


def build(R, kw, seed, plan, extra):
    gcfg = [make_config(**kw) for _ in range(R)]
    ocfg = [O.make_config(**kw) for _ in range(R)]
    engines = [Engine(gcfg[r], seed=seed + r, device=0, region=r, plan=plan, extra_capacity=extra) for r in range(R)]
    orc = O.OracleMultiEngine(ocfg, seed=seed, migration=plan.get("migration"), commute=plan.get("commute"),
                              start_migration_hour=plan.get("start_migration_hour", 0), end_migration_hour=plan.get("end_migration_hour", 0),
                              extra_capacity=extra, threads=2)
    return engines, orc


def assert_regions_equal(engines, orc, ctx):
    for r, e in enumerate(engines):
        a, b = e.get_state(), orc.get_state(r)
        assert e.population == orc.population(r), f"{ctx}: region {r} population {e.population} vs {orc.population(r)}"
        for f in STATE_FIELDS:
            bad = np.nonzero(a[f] != b[f])[0]
            assert bad.size == 0, f"{ctx}: region {r} field {f} differs for {bad.size} slots, first {bad[:5]}: gpu {a[f][bad[:5]]} oracle {b[f][bad[:5]]}"
        reg = e.get_regions()
        bad = np.nonzero(reg != b["reg"])[0]
        assert bad.size == 0, f"{ctx}: region {r} reg differs for {bad.size} slots, first {bad[:5]}: gpu {reg[bad[:5]]} oracle {b['reg'][bad[:5]]}"


def test_three_regions_commute_and_migration_bit_exact():
    R = 3
    kw = dict(n_agents=3000, grid_size=200, hours=400, exposed=60, asym=10, mild=10, severe=10)
    plan = dict(n_regions=R, migration=np.array([[0, 40, 20], [30, 0, 10], [25, 15, 0]], np.uint32),
                commute=np.array([[0, 30, 10], [20, 0, 15], [5, 25, 0]], np.uint32), start_migration_hour=20, end_migration_hour=150)
    engines, orc = build(R, kw, 31, plan, 600)
    try:
        assert_regions_equal(engines, orc, "init")
        m = MultiRegion(engines)
        hour = 1
        for day in range(7):
            rows = m.run(hour, 24)
            for k in range(24):
                want = orc.step(hour + k)
                for r in range(R):
                    assert (rows[r, k] == want[r]).all(), f"hour {hour + k} region {r}: gpu {rows[r, k]} oracle {want[r]}"
            hour += 24
            assert_regions_equal(engines, orc, f"end of day {day}")
        assert sum(e.population for e in engines) == R * 3000
    finally:
        for e in engines:
            e.close()


def test_commute_only_hour_by_hour_state():
    R = 2
    kw = dict(n_agents=4000, grid_size=220, hours=200, exposed=100, mild=20, severe=20, lockdown=(30, 0.2))
    plan = dict(n_regions=R, commute=np.array([[0, 120], [80, 0]], np.uint32))
    engines, orc = build(R, kw, 77, plan, 400)
    try:
        m = MultiRegion(engines)
        for hour in range(1, 24 * 3 + 1):
            rows = m.run(hour, 1)
            want = orc.step(hour)
            for r in range(R):
                assert (rows[r, 0] == want[r]).all(), f"hour {hour} region {r}: gpu {rows[r, 0]} oracle {want[r]}"
            if hour % 24 in (7, 8, 16, 17, 18, 0):
                assert_regions_equal(engines, orc, f"hour {hour}")
        for r in range(R):  # the lockdown fired in both
            ev = engines[r].intervention_events()
            assert (ev[:, 1] == 0).any()
            assert [tuple(int(v) for v in x) for x in ev] == [tuple(int(v) for v in x) for x in orc.events(r)]
    finally:
        for e in engines:
            e.close()


def test_out_of_slots_is_an_error_not_a_crash():
    from epirust_b200.engine import EpiError

    R = 2
    kw = dict(n_agents=2000, grid_size=160, hours=100, exposed=10)
    plan = dict(n_regions=R, commute=np.array([[0, 50], [0, 0]], np.uint32))
    engines, _ = build(R, kw, 5, plan, 10)
    try:
        m = MultiRegion(engines)
        with pytest.raises(EpiError, match="out of agent slots"):
            m.run(1, 8)
    finally:
        for e in engines:
            e.close()


def test_segment_overflow_is_reported_and_leaves_the_region_intact():
    """epi_travel_pack into caller-owned segments that are too small: the error is reported and nobody was removed
    (k_travel_pack used to vacate the leavers before the overflow was known)."""
    import torch

    from epirust_b200 import _ffi
    from epirust_b200.engine import EpiError

    R = 2
    kw = dict(n_agents=2000, grid_size=160, hours=100, exposed=10)
    plan = dict(n_regions=R, commute=np.array([[0, 50], [0, 0]], np.uint32))
    engines, _ = build(R, kw, 5, plan, 200)
    try:
        e = engines[0]
        for h in range(1, 8):
            e.step(h)
        before = e.get_state()
        send = torch.zeros((R, 16, 8), dtype=torch.int32, device="cuda")  # 50 commuters do not fit a 15-record segment
        with pytest.raises(EpiError, match="stride_records"):
            e.travel_pack(7, _ffi.TRAVEL_COMMUTE, send.data_ptr(), 16)
        assert e.population == 2000
        after = e.get_state()
        for f in STATE_FIELDS:
            assert (before[f] == after[f]).all(), f
        assert int(send[:, 0, 0].sum()) == 0  # the headers promise no records
        big = torch.zeros((R, 64, 8), dtype=torch.int32, device="cuda")
        counts = e.travel_pack(7, _ffi.TRAVEL_COMMUTE, big.data_ptr(), 64)
        assert counts.tolist() == [0, 50] and e.population == 1950
    finally:
        for e in engines:
            e.close()


def test_crowded_housing_needs_more_than_three_placement_rounds():
    """select_starting_points (allocation_map.rs:339-347) in a housing strip that is ~97 % full at midnight: some arrival finds all
    24 candidates of the first three rounds taken, so the placement kernel (k_travel_place) must loop beyond three rounds."""
    R = 2
    kw = dict(n_agents=1420, grid_size=60, hours=200, exposed=20, pt=0.0, working=0.2)
    plan = dict(n_regions=R, migration=np.array([[0, 15], [15, 0]], np.uint32), start_migration_hour=20, end_migration_hour=150)
    engines, orc = build(R, kw, 3, plan, 1000)
    try:
        m = MultiRegion(engines)
        hour = 1
        for day in range(4):
            rows = m.run(hour, 24)
            for k in range(24):
                want = orc.step(hour + k)
                for r in range(R):
                    assert (rows[r, k] == want[r]).all(), f"hour {hour + k} region {r}: gpu {rows[r, k]} oracle {want[r]}"
            hour += 24
            assert_regions_equal(engines, orc, f"end of day {day}")
        assert max(orc.max_place_rounds(r) for r in range(R)) > 3
    finally:
        for e in engines:
            e.close()


def test_many_migrators_fill_houses_level_by_level():
    """A heavy migration plan: thousands of arrivals per exchange exercise the parallel water filling of the occupancy heaps
    over several occupancy levels (houses hold at most 4), against the oracle's sequential BinaryHeap."""
    R = 2
    kw = dict(n_agents=6000, grid_size=200, hours=200, exposed=50, working=0.6)
    plan = dict(n_regions=R, migration=np.array([[0, 2500], [100, 0]], np.uint32), start_migration_hour=10, end_migration_hour=60)
    engines, orc = build(R, kw, 13, plan, 6000)
    try:
        m = MultiRegion(engines)
        hour = 1
        for day in range(3):
            rows = m.run(hour, 24)
            for k in range(24):
                want = orc.step(hour + k)
                for r in range(R):
                    assert (rows[r, k] == want[r]).all(), f"hour {hour + k} region {r}: gpu {rows[r, k]} oracle {want[r]}"
            hour += 24
            assert_regions_equal(engines, orc, f"end of day {day}")
        assert engines[1].population > 6000 + 4000  # two exchanges of ~2500 arrivals: the 2000 one-resident houses fill first, then level 2
    finally:
        for e in engines:
            e.close()
