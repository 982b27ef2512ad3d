"""The optional tile kernels of the plain movement hours (csrc/tiles.cu: one warp per run of adjacent offices / houses, grid
bytes staged by TMA, lowest-id claims settled on chip; epi_set_tiles) against the oracle and against the global-path kernels (epi_set_tiles(0)): bit-exact, also on
crafted states that break the tiles' assumptions (agents whose current_area is the whole housing strip in the evening,
public-transport commuters at home, agents standing in other people's houses / outside their office, hospital staff that claim
an office) -- those must be caught by the dirty marks / the housing count and settled on the global path."""
import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200.engine import Engine, make_config, STATE_FIELDS

pytestmark = pytest.mark.gpu

WS_SHIFT, AREA_SHIFT = 13, 15
WS_NORMAL, WS_ESSENTIAL, WS_STAFF, WS_NA = 0, 1, 2, 3
AK_HOME, AK_WORK, AK_TRANSPORT, AK_HOUSING = 0, 1, 2, 3
ST_PT = 1 << 9


def assert_state_equal(a, b, ctx):
    for f in STATE_FIELDS:
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, f"{ctx}: field {f} differs for {bad.size} agents, first {bad[:5]}: tiles {a[f][bad[:5]]} other {b[f][bad[:5]]}"


def test_tiles_equal_the_global_path_over_ten_days():
    kw = dict(n_agents=300_000, grid_size=1370, hours=400, exposed=3000, asym=300, mild=300, severe=300, lockdown=(20_000, 0.1), vaccinate=((60, 0.2),))
    with Engine(make_config(**kw), seed=3) as tiles, Engine(make_config(**kw), seed=3) as plain:
        tiles.set_tiles(True)
        for day in range(10):
            a, _ = tiles.simulate_hours(24 * day + 1, 24)
            b, _ = plain.simulate_hours(24 * day + 1, 24)
            assert (a == b).all(), f"day {day}: first differing hour {24 * day + 1 + int(np.nonzero((a != b).any(axis=1))[0][0])}"
            assert_state_equal(tiles.get_state(), plain.get_state(), f"end of day {day}")
        assert tiles.tile_hours == 11 * 10 and plain.tile_hours == 0
        assert (tiles.get_grid() == plain.get_grid()).all()
        # hour by hour through the evening and the office hours of the next day (single steps: no graph)
        for hour in range(241, 241 + 48):
            assert (tiles.step(hour) == plain.step(hour)).all(), f"hour {hour}"
            if hour % 24 in (9, 11, 13, 15, 18, 22):
                assert_state_equal(tiles.get_state(), plain.get_state(), f"hour {hour}")
        # switching the tiles off and on again in mid-run changes nothing
        tiles.set_tiles(False)
        assert (tiles.step(289) == plain.step(289)).all()
        tiles.set_tiles(True)
        a, _ = tiles.simulate_hours(290, 47)
        b, _ = plain.simulate_hours(290, 47)
        assert (a == b).all()
        assert_state_equal(tiles.get_state(), plain.get_state(), "after the toggle")


def run_against_oracle(kw, seed, craft, base_hour, hours):
    """advance to base_hour - 1, apply craft(state) to both, then compare every hour's Counts and full state"""
    with Engine(make_config(**kw), seed=seed) as gpu:
        gpu.set_tiles(True)
        orc = O.OracleEngine(O.make_config(**kw), seed=seed)
        for h in range(1, base_hour):
            assert (gpu.step(h) == orc.step(h)).all(), f"hour {h}"
        s = gpu.get_state()
        craft(s, gpu)
        gpu.set_state(s), orc.set_state(s)
        assert_state_equal(gpu.get_state(), orc.get_state(), "crafted state")
        t0 = gpu.tile_hours
        for hour in range(base_hour, base_hour + hours):
            cg, co = gpu.step(hour), orc.step(hour)
            assert (cg == co).all(), f"hour {hour}: gpu {cg} oracle {co}"
            assert_state_equal(gpu.get_state(), orc.get_state(), f"hour {hour}")
        assert gpu.tile_hours > t0
        return gpu.tile_hours - t0


KW = dict(n_agents=8000, grid_size=170, hours=2000, exposed=300, asym=60, mild=60, severe=60)


def test_evening_with_housing_strip_walkers_and_commuters_at_home():
    def craft(s, gpu):
        st = s["st"]
        ws = (st >> WS_SHIFT) & 3
        kind = (st >> AREA_SHIFT) & 7
        na = np.nonzero(ws == WS_NA)[0]
        # 40 non-working agents roam the whole housing strip (they may step into anybody's house: the house tiles must stand down)
        st[na[:40]] = (st[na[:40]] & ~np.uint32(7 << AREA_SHIFT)) | np.uint32(AK_HOUSING << AREA_SHIFT)
        # public-transport commuters whose current_area is their home although they are in the generic segment (dirty marks)
        pt = np.nonzero(((ws == WS_NORMAL) | (ws == WS_ESSENTIAL)) & ((st & ST_PT) != 0))[0]
        st[pt[:200]] = (st[pt[:200]] & ~np.uint32(7 << AREA_SHIFT)) | np.uint32(AK_HOME << AREA_SHIFT)
        assert (kind[pt[:200]] == AK_TRANSPORT).any()

    run_against_oracle(KW, 21, craft, 24 * 2 + 19, 30)


def test_evening_with_agents_in_foreign_houses_and_commuters_at_home():
    def craft(s, gpu):
        st = s["st"]
        ws = (st >> WS_SHIFT) & 3
        pt = np.nonzero(((ws == WS_NORMAL) | (ws == WS_ESSENTIAL)) & ((st & ST_PT) != 0))[0]
        st[pt[:300]] = (st[pt[:300]] & ~np.uint32(7 << AREA_SHIFT)) | np.uint32(AK_HOME << AREA_SHIFT)  # generic agents that walk at home
        # swap the positions of pairs of agents that are at home: each now stands in the other's house with current_area = own home
        kind = (st >> AREA_SHIFT) & 7
        home = np.nonzero((kind == AK_HOME) & (ws != WS_STAFF))[0]
        geo = gpu.geometry()
        in_housing = (s["cell_x"][home] <= geo[2]) & (s["cell_y"][home] < geo[18])
        home = home[in_housing][:600]
        a, b = home[0::2], home[1::2]
        n = min(len(a), len(b))
        a, b = a[:n], b[:n]
        for f in ("cell_x", "cell_y"):
            s[f][a], s[f][b] = s[f][b].copy(), s[f][a].copy()

    run_against_oracle(KW, 22, craft, 24 * 3 + 18, 29)


def test_office_hours_with_outsiders():
    def craft(s, gpu):
        st = s["st"]
        ws = (st >> WS_SHIFT) & 3
        workers = np.nonzero((ws == WS_NORMAL) | (ws == WS_ESSENTIAL))[0]
        # workers swap places pairwise: everybody stands in somebody else's office (or at home) with current_area = own office
        a, b = workers[0:800:2], workers[1:800:2]
        for f in ("cell_x", "cell_y"):
            s[f][a], s[f][b] = s[f][b].copy(), s[f][a].copy()
        # hospital staff that believe they work in an office (generic segment -> dirty office tiles)
        staff = np.nonzero(ws == WS_STAFF)[0]
        st[staff] = (st[staff] & ~np.uint32(7 << AREA_SHIFT)) | np.uint32(AK_WORK << AREA_SHIFT)
        s["wsa"][staff] = 5000  # not on duty: they walk
        extra = workers[800:1000]  # ... and some more staff, made from workers
        st[extra] = (st[extra] & ~np.uint32(3 << WS_SHIFT)) | np.uint32(WS_STAFF << WS_SHIFT)
        st[extra] = (st[extra] & ~np.uint32(7 << AREA_SHIFT)) | np.uint32(AK_WORK << AREA_SHIFT)
        s["wsa"][extra] = 5000

    run_against_oracle(KW, 23, craft, 24 * 2 + 9, 16)


def test_crowded_offices_one_office_per_tile(monkeypatch):
    """offices with more workers than cells (the losers keep trying from outside), one office per tile"""
    monkeypatch.setenv("EPI_TILE_OFFICES", "1")
    kw = dict(n_agents=3600, grid_size=100, hours=200, exposed=300)  # 20 offices with ~126 workers each
    with Engine(make_config(**kw), seed=4) as gpu:
        gpu.set_tiles(True)
        orc = O.OracleEngine(O.make_config(**kw), seed=4)
        for hour in range(1, 49):
            assert (gpu.step(hour) == orc.step(hour)).all(), f"hour {hour}"
        assert_state_equal(gpu.get_state(), orc.get_state(), "end")
        assert gpu.tile_hours == 22
