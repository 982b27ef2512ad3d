"""The CPU oracle against every known-answer value the reference's own unit tests pin for the hot path
(values transcribed in tests/golden/reference_kats.json with file:line citations)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_ffi as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K = json.load(open(os.path.join(GOLDEN, "reference_kats.json")))


@pytest.fixture(scope="module")
def L():
    return O.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_philox_known_answers(L):
    for v in json.load(open(os.path.join(GOLDEN, "philox_kat.json")))["vectors"]:
        ctr = np.array([int(x, 16) for x in v["ctr"]], np.uint32)
        key = np.array([int(x, 16) for x in v["key"]], np.uint32)
        out = np.zeros(4, np.uint32)
        L.orc_kat_philox(_p(ctr), _p(key), _p(out))
        assert [f"{x:08x}" for x in out] == v["out"]


def test_draw_slot_convention(L):
    seed, agent, hour = 0x1234_5678_9ABC_DEF0, 77, 31
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], np.uint32)

    def block(b, dom):
        out = np.zeros(4, np.uint32)
        L.orc_kat_philox(_p(np.array([agent, hour, b, dom], np.uint32)), _p(key), _p(out))
        return [int(v) for v in out]

    # generic domains: u64 slot s lives in Philox block s>>1 of counter (agent, hour, s>>1, domain); even = words 0,1; odd = words 2,3
    for slot in range(16):
        o = block(slot >> 1, 1)
        lo, hi = (o[2], o[3]) if slot & 1 else (o[0], o[1])
        assert L.orc_kat_draw(seed, agent, hour, 1, slot) == lo | (hi << 32)
    # hour-step domain: block 0 = PICK, FACTOR, A ; block 1 = PX, PY ; block 2+(j>>1) = EXPOSE j
    o0, o1 = block(0, 0), block(1, 0)
    assert L.orc_kat_draw32(seed, agent, hour, 0) == o0[0]  # SLOT_PICK
    assert L.orc_kat_draw32(seed, agent, hour, 1) == o0[1]  # SLOT_FACTOR
    assert L.orc_kat_draw(seed, agent, hour, 0, 2) == o0[2] | (o0[3] << 32)  # SLOT_A
    assert L.orc_kat_draw32(seed, agent, hour, 3) == o1[0]  # SLOT_PX
    assert L.orc_kat_draw32(seed, agent, hour, 4) == o1[1]  # SLOT_PY
    for j in range(8):
        o = block(2 + (j >> 1), 0)
        lo, hi = (o[2], o[3]) if j & 1 else (o[0], o[1])
        assert L.orc_kat_draw(seed, agent, hour, 0, 8 + j) == lo | (hi << 32)


def test_bernoulli_threshold_is_rand_0_8(L):
    # rand 0.8 Bernoulli::new: p_int = (p * 2^64) as u64; p == 1.0 is ALWAYS_TRUE
    assert L.orc_kat_bernoulli_threshold(1.0) == 2**64 - 1
    assert L.orc_kat_bernoulli_threshold(0.0) == 0
    assert L.orc_kat_bernoulli_threshold(0.5) == 2**63
    assert L.orc_kat_bernoulli_threshold(0.25) == 2**62
    assert L.orc_kat_bernoulli_threshold(0.035) == int(0.035 * 2.0**64)


def test_neighbor_order(L):
    k = K["neighbor_order"]
    out = np.zeros(16, np.int32)
    L.orc_kat_neighbors(k["point"][0], k["point"][1], _p(out))
    assert out.reshape(8, 2).tolist() == k["expect"]


def test_area_neighbors_inclusive_bounds(L):
    k = K["area_neighbors"]
    out = np.zeros(16, np.int32)
    n = L.orc_kat_area_neighbors(*k["area"], *k["point"], _p(out))
    pts = out.reshape(8, 2)[:n].tolist()
    assert n == k["count"] and k["contains"] in pts and k["not_contains"] not in pts


def test_area_iter_order(L):
    for case in K["area_iter"]["cases"]:
        out = np.zeros(2 * 64, np.int32)
        n = L.orc_kat_area_iter(*case["area"], _p(out), 64)
        assert out[: 2 * n].reshape(n, 2).tolist() == case["expect"]


def test_area_factory(L):
    k = K["area_factory"]
    out = np.zeros(4 * 64, np.int32)
    n = L.orc_kat_area_factory(*k["start"], *k["end"], k["size"], _p(out), 64)
    assert n == k["count"]
    areas = out[: 4 * n].reshape(n, 4).tolist()
    for idx, rect in k["areas"].items():
        assert areas[int(idx)] == rect
    for idx, pt, want in k["contains"]:
        assert bool(L.orc_kat_area_contains(*areas[idx], *pt)) == want


def test_area_contains_and_number_of_cells(L):
    k = K["area_contains"]
    assert L.orc_kat_area_contains(*k["area"], *k["inside"]) == 1
    assert L.orc_kat_area_contains(*k["area"], *k["outside"]) == 0
    k = K["number_of_cells"]
    assert L.orc_kat_number_of_cells(*k["area"]) == k["expect"]  # (ex-sx)*(ey-sy): 25, not 36


def test_define_geography(L):
    k = K["define_geography"]
    out = np.zeros(19, np.int32)
    L.orc_kat_define_geography(k["grid_size"], _p(out))
    assert out[0:4].tolist() == k["housing"] and out[4:8].tolist() == k["transport"]
    assert out[8:12].tolist() == k["work"] and out[12:16].tolist() == k["hospital"]


def test_resize_and_increase_hospital(L):
    k = K["resize_hospital"]
    for case in k["cases"]:
        out = np.zeros(19, np.int32)
        L.orc_kat_resize_hospital(k["grid_size"], case["agents"], case["staff"], case["beds"], _p(out))
        assert out[12:16].tolist() == case["hospital"]
    k = K["increase_hospital"]
    out = np.zeros(19, np.int32)
    L.orc_kat_increase_hospital(k["grid_size"], k["new_size"], _p(out))
    assert out[12:16].tolist() == k["hospital"]


def test_goto_hospital(L):
    k = K["goto_hospital"]
    occ = np.array(k["occupied"], np.int32)
    out = np.zeros(2, np.int32)
    r = L.orc_kat_goto_hospital(k["grid_size"], _p(occ), len(occ), *k["hospital"], *k["home"], *k["cell"], 1, _p(out))
    assert bool(r) == k["hospitalized"] and out.tolist() == k["new_cell"]
    k = K["goto_hospital_full"]
    occ = np.array(k["occupied"], np.int32)
    for seed in range(20):
        r = L.orc_kat_goto_hospital(k["grid_size"], _p(occ), len(occ), *k["hospital"], *k["home"], *k["cell"], seed, _p(out))
        assert not r
        assert L.orc_kat_area_contains(*k["home"], int(out[0]), int(out[1]))


def test_is_point_in_grid(L):
    k = K["point_in_grid"]
    for x, y in k["inside"]:
        assert L.orc_kat_is_point_in_grid(k["grid_size"], x, y)
    for x, y in k["outside"]:
        assert not L.orc_kat_is_point_in_grid(k["grid_size"], x, y)


def test_small_pox_rates(L):
    k = K["small_pox"]
    cfg = O.make_config(**k["disease"])
    for day, rate in k["rate"].items():
        assert L.orc_kat_transmission_rate(C.byref(cfg), int(day)) == rate
    for day, want in k["hospitalized"].items():
        assert bool(L.orc_kat_is_to_be_hospitalized(C.byref(cfg), int(day))) == want
    # u32 wrap of (day + immunity) < 0: rate 0 (citizen/mod.rs:182-185)
    assert L.orc_kat_transmission_rate(C.byref(cfg), (0 - 2) & 0xFFFFFFFF) == 0.0


def _counts(hour, s, e, i, h, r, d):
    return np.array([hour, s, e, i, h, r, d], np.uint32)


def test_lockdown_gating(L):
    k = K["lockdown"]
    cfg = O.make_config(lockdown=(k["at_number_of_infections"], k["essential_workers_population"]))
    op = lambda iv, o, c, arg=0: L.orc_iv_op(iv, o, _p(c), arg)
    iv = L.orc_iv_create(C.byref(cfg))
    # should_apply_lockdown_at_threshold
    assert not op(iv, 0, _counts(0, 99, 0, 1, 0, 0, 0))
    assert not op(iv, 0, _counts(22, 80, 0, 20, 0, 0, 0))
    assert not op(iv, 0, _counts(28, 79, 0, 21, 0, 0, 0))
    assert op(iv, 0, _counts(48, 79, 0, 21, 0, 0, 0))
    assert op(iv, 1, _counts(0, 0, 0, 0, 0, 0, 0)) == 1
    # should_not_apply_lockdown_when_already_locked_down
    assert not op(iv, 0, _counts(48, 75, 0, 25, 0, 0, 0))
    # should_lift_lockdown_at_after_time_elapsed...
    until = 48 + 7 * 24
    op(iv, 4, _counts(0, 0, 0, 0, 0, 0, 0), until)
    for hr in range(48, until):
        assert not op(iv, 2, _counts(hr, 80, 0, 20, 0, 0, 0))
    assert not op(iv, 2, _counts(until, 79, 0, 21, 0, 0, 0))
    assert not op(iv, 2, _counts(until + 1, 79, 0, 20, 0, 0, 0))
    assert op(iv, 2, _counts(until + 21 * 24, 80, 0, 20, 0, 0, 0))
    L.orc_iv_destroy(iv)
    # should_not_reapply_lockdown
    iv = L.orc_iv_create(C.byref(cfg))
    op(iv, 1, _counts(0, 0, 0, 0, 0, 0, 0))
    op(iv, 4, _counts(0, 0, 0, 0, 0, 0, 0), 28)
    assert op(iv, 2, _counts(532, 80, 0, 20, 0, 0, 0))
    assert not op(iv, 0, _counts(540, 70, 0, 30, 0, 0, 0))
    L.orc_iv_destroy(iv)


def test_hospital_intervention_gating(L):
    k = K["hospital_intervention"]
    op = lambda iv, o, c: L.orc_iv_op(iv, o, _p(c), 0)
    c0 = _counts(0, 99, 1, 0, 0, 0, 0)
    iv = L.orc_iv_create(C.byref(O.make_config(hospital=k["spread_rate_threshold"])))
    op(iv, 5, c0)
    assert not op(iv, 6, c0)
    op(iv, 5, _counts(24, 80, 0, 20, 0, 0, 0))
    assert op(iv, 6, c0)
    L.orc_iv_destroy(iv)
    iv = L.orc_iv_create(C.byref(O.make_config()))  # intervention absent
    op(iv, 5, c0)
    op(iv, 5, _counts(24, 80, 0, 20, 0, 0, 0))
    assert not op(iv, 6, c0)
    L.orc_iv_destroy(iv)
    iv = L.orc_iv_create(C.byref(O.make_config(hospital=k["spread_rate_threshold"])))  # below threshold
    op(iv, 5, c0)
    op(iv, 5, _counts(24, 95, 0, 5, 0, 0, 0))
    assert not op(iv, 6, c0)
    L.orc_iv_destroy(iv)
    iv = L.orc_iv_create(C.byref(O.make_config(hospital=k["spread_rate_threshold"])))  # already applied
    assert op(iv, 7, c0) == 1
    op(iv, 5, _counts(24, 80, 0, 20, 0, 0, 0))
    assert not op(iv, 6, c0)
    L.orc_iv_destroy(iv)


def test_vaccination_lookup(L):
    k = K["vaccination"]
    iv = L.orc_iv_create(C.byref(O.make_config(vaccinate=((k["at_hour"], k["percent"]),))))
    assert L.orc_iv_op(iv, 8, _p(_counts(5000, 10, 0, 10, 10, 10, 10)), 0) == int(round(k["percent"] * 1e6))
    assert L.orc_iv_op(iv, 8, _p(_counts(5001, 10, 0, 10, 10, 10, 10)), 0) == -1
    L.orc_iv_destroy(iv)


def test_counts_update(L):
    # Counts::update_counts (counts.rs:126-140): Infected && hospitalized -> hospitalized column
    st = np.array([0, 0, 1, 2, 2 | (1 << 10), 3, 4, 4], np.uint32)
    out = np.zeros(7, np.uint32)
    L.orc_kat_counts(_p(st), len(st), _p(out))
    assert out.tolist() == [0, 2, 1, 1, 1, 1, 2]


def test_starting_infections_by_category():
    # citizen_factory.rs:190-215: the configured numbers of each category are assigned, the rest stay susceptible
    cfg = O.make_config(n_agents=2000, grid_size=100, exposed=5, asym=2, mild=3, severe=4)
    e = O.OracleEngine(cfg, seed=3)
    st = e.get_state()["st"]
    state, sev, day = st & 7, (st >> 3) & 3, st >> 18
    assert (state == 1).sum() == 5
    assert ((state == 2) & (sev == 1)).sum() == 2 and ((state == 2) & (sev == 2)).sum() == 3 and ((state == 2) & (sev == 3)).sum() == 4
    assert (day[state == 2] == 1).all()
    assert (state == 0).sum() == 2000 - 14
    assert e.counts_at_start().tolist() == [0, 1986, 5, 9, 0, 0, 0]


def test_population_invariants():
    # grid.rs:352-376: homes inside the housing strip, offices inside the work strip, distinct start cells
    cfg = O.make_config(n_agents=5000, grid_size=150)
    e = O.OracleEngine(cfg, seed=2)
    s, g = e.get_state(), e.geometry()
    assert len(set(zip(s["cell_x"].tolist(), s["cell_y"].tolist()))) == 5000
    assert (s["cell_x"] >= g[0]).all() and (s["cell_x"] <= g[2]).all()
    ws = (s["st"] >> 13) & 3
    assert 0.6 < (ws != 3).mean() < 0.8  # working_percentage 0.7
    from test_population_host import check_numbering

    check_numbering(s, 5000, int(g[16]), int(g[17]))


def test_oracle_default_json_run_is_deterministic_and_conserves_population():
    cfg = O.default_json_config()
    rows_a, _, _ = O.oracle_run(cfg, seed=5, mode="keyed", threads=2, max_hours=200)
    rows_b, _, _ = O.oracle_run(cfg, seed=5, mode="keyed", threads=4, max_hours=200)
    assert (rows_a == rows_b).all()  # keyed draws: independent of the thread count
    assert (rows_a[:, 1:].sum(axis=1) == 10000).all()  # allocation_map.rs:128
    assert (rows_a[:, 0] == np.arange(1, len(rows_a) + 1)).all()
