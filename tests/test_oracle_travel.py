"""CPU tests of the multi-region oracle: the reference's pinned travel-plan values (engine_migration_plan.rs:120-150,
common/src/models/migration_plan.rs:50-75) and invariants of the restated exchange."""
import ctypes as C

import numpy as np

import oracle_ffi as O

MATRIX = np.array([[0, 156, 24], [108, 0, 221], [97, 12, 0]], np.uint32)  # migration_plan.rs:77-80


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_percent_outgoing_and_allocation_known_answers():
    L = O._multi_lib()
    assert L.orc_kat_percent_outgoing(_p(MATRIX), 3, 0, 10000) == 0.018  # engine_migration_plan.rs:121-124
    out = np.zeros(3, np.uint32)
    L.orc_kat_alloc_outgoing(_p(MATRIX), 3, 0, 180, _p(out))  # :136-150: 180 outgoing -> 156 / 24
    assert out.tolist() == [0, 156, 24]
    # row sums (migration_plan.rs:52-57)
    assert [int(MATRIX[r].sum()) for r in range(3)] == [156 + 24, 108 + 221, 97 + 12]
    # fewer travellers than planned: proportional shares, floor, surplus stays (migrators_by_engine.rs:27-32)
    L.orc_kat_alloc_outgoing(_p(MATRIX), 3, 0, 147, _p(out))
    assert out.tolist() == [0, int(156 / 180 * 147), int(24 / 180 * 147)]


def _cfgs(R, n=3000, g=200, **kw):
    base = dict(n_agents=n, grid_size=g, hours=200, exposed=40, asym=5, mild=5, severe=5)
    base.update(kw)
    return [O.make_config(**base) for _ in range(R)]


def test_exchange_conserves_agents_and_counts():
    R = 3
    mig = np.array([[0, 40, 20], [30, 0, 10], [25, 15, 0]], np.uint32)
    com = np.array([[0, 30, 10], [20, 0, 15], [5, 25, 0]], np.uint32)
    # no symptomatic agents on day 1, so every planned commuter can_move (citizen/mod.rs:452-454, 488-495)
    m = O.OracleMultiEngine(_cfgs(R, mild=0, severe=0), seed=5, migration=mig, commute=com, start_migration_hour=20, end_migration_hour=150, extra_capacity=600, threads=2)
    total0 = sum(m.population(r) for r in range(R))
    pops = []
    for hour in range(1, 24 * 4 + 1):
        rows = m.step(hour)
        for r in range(R):
            assert rows[r, 1:].sum() == m.population(r), f"hour {hour} region {r}"
        assert sum(m.population(r) for r in range(R)) == total0  # nobody is lost or duplicated in transit
        pops.append([m.population(r) for r in range(R)])
    pops = np.array(pops)
    # commuters are away between 07:00 and 17:00: region 0 sends 40 and hosts 25
    assert pops[7 - 1, 0] == 3000 - 40 + 25 and pops[17 - 1, 0] == 3000
    # migration happens at hours 24, 48, 72 (start 20 < h < end 150) and changes the resident populations
    assert (pops[24 - 1] != 3000).any()
    st = m.get_state(0)
    alive = (st["st"] & 7) != 7
    assert alive.sum() == m.population(0)
    # visitors keep their foreign home region; residents that commute out carry a foreign work region
    assert ((st["reg"][alive] & 0xFF) == 0).all()  # at hour 96 (h = 0) every visitor has gone home
    assert (((st["reg"][alive] >> 8) & 0xFF) != 0).sum() >= 40


def test_commuters_become_normal_workers_of_the_host_region():
    R = 2
    com = np.array([[0, 50], [0, 0]], np.uint32)
    m = O.OracleMultiEngine(_cfgs(R), seed=9, commute=com, extra_capacity=100)
    for hour in range(1, 9):
        m.step(hour)
    st = m.get_state(1)
    visitors = ((st["st"] & 7) != 7) & ((st["reg"] & 0xFF) == 0)
    assert visitors.sum() == 50 and m.population(0) == 3000 - 50
    ws = (st["st"][visitors] >> 13) & 3
    assert (ws == 0).all()  # WorkStatus::Normal (citizen/mod.rs:138-154)
    assert (((st["reg"][visitors] >> 8) & 0xFF) == 1).all()  # office assigned in the host region at hour 7 (allocation_map.rs:260-268)
    iso = (st["st"][visitors] >> 11) & 1
    assert (iso == 0).all()
