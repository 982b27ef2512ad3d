"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact -- all arithmetic on the path is integer, plus f64 probabilities turned into u64 thresholds on the host.
"""
import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200.engine import Engine, EpiError, make_config, run_standalone, STATE_FIELDS

pytestmark = pytest.mark.gpu


def assert_state_equal(gpu, orc, ctx=""):
    a, b = gpu.get_state(), orc.get_state()
    for f in STATE_FIELDS:
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, f"{ctx}: field {f} differs for {bad.size} agents, first {bad[:5]}: gpu {a[f][bad[:5]]} oracle {b[f][bad[:5]]}"


def small_cfg(**kw):
    base = dict(n_agents=4000, grid_size=120, hours=400, exposed=30, asym=5, mild=5, severe=5)
    base.update(kw)
    return make_config(**base), O.make_config(**base)


def test_initial_state_and_geometry_match_oracle():
    for n, g in ((4000, 120), (10000, 250), (777, 60)):
        gc, oc = small_cfg(n_agents=n, grid_size=g)
        with Engine(gc, seed=11) as gpu:
            orc = O.OracleEngine(oc, seed=11)
            assert (gpu.geometry() == orc.geometry()).all()
            assert (gpu.counts_at_start() == orc.counts_at_start()).all()
            assert_state_equal(gpu, orc, f"init n={n}")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_hour_by_hour_bit_exact(seed):
    gc, oc = small_cfg()
    with Engine(gc, seed=seed) as gpu:
        orc = O.OracleEngine(oc, seed=seed)
        for hour in range(1, 24 * 14 + 1):
            cg, co = gpu.step(hour), orc.step(hour)
            assert (cg == co).all(), f"hour {hour}: gpu {cg} oracle {co}"
            if hour % 24 in (0, 7, 8, 12, 16, 17, 23) or hour < 30:
                assert_state_equal(gpu, orc, f"hour {hour}")


def test_dense_epidemic_with_fast_disease():
    # high rates and short durations so every transition (S->E->I->R/D, hospitalisation, hospital-full) happens often
    kw = dict(n_agents=3600, grid_size=100, exposed=200, severe=100, mild=50, asym=50, beds=0.0005,
              regular_transmission_rate=0.6, high_transmission_rate=0.9, death_rate=0.4, exposed_duration=10, pre_symptomatic_duration=8,
              last_day=6, regular_transmission_start_day=1, high_transmission_start_day=3, percentage_severe_infected_population=0.6)
    gc, oc = small_cfg(**kw)
    with Engine(gc, seed=5) as gpu:
        orc = O.OracleEngine(oc, seed=5)
        for hour in range(1, 24 * 12 + 1):
            cg, co = gpu.step(hour), orc.step(hour)
            assert (cg == co).all(), f"hour {hour}: gpu {cg} oracle {co}"
            if hour % 24 in (0, 23, 8):
                assert_state_equal(gpu, orc, f"hour {hour}")
        assert cg[6] > 0 and cg[5] > 0  # deaths and recoveries happened
        assert_state_equal(gpu, orc, "end")


@pytest.mark.parametrize("hour", [24, 31, 32, 36, 40, 41, 47, 48, 55])
def test_substep_with_injected_draws(hour):
    """Deterministic sub-steps given identical injected draws (BASELINE.json north_star)."""
    gc, oc = small_cfg(exposed=300, severe=100, mild=100, asym=100)
    rng = np.random.default_rng(hour)
    with Engine(gc, seed=9) as gpu:
        orc = O.OracleEngine(oc, seed=9)
        # advance both to `hour` with keyed draws so the state is a realistic mid-run state
        for h in range(1, hour):
            gpu.step(h), orc.step(h)
        draws = rng.integers(0, 2**64, size=(gpu.population, 16), dtype=np.uint64)
        # make Bernoulli successes common so transitions fire
        draws[:, 2] >>= np.uint64(rng.integers(0, 3))  # SLOT_A (u64 Bernoulli draw)
        draws[:, 8:16] >>= np.uint64(2)  # SLOT_EXPOSE0..7
        cg, co = gpu.step(hour, draws), orc.step(hour, draws)
        assert (cg == co).all()
        assert_state_equal(gpu, orc, f"injected hour {hour}")


def test_intervention_sweeps_match():
    gc, oc = small_cfg(lockdown=(10, 0.2))
    with Engine(gc, seed=4) as gpu:
        orc = O.OracleEngine(oc, seed=4)
        for h in range(1, 25):
            gpu.step(h), orc.step(h)
        gpu.lock_city(), orc.lock_city()
        assert_state_equal(gpu, orc, "lock")
        gpu.vaccinate(0.35, 24), orc.vaccinate(0.35, 24)
        assert_state_equal(gpu, orc, "vaccinate")
        for h in range(25, 49):
            assert (gpu.step(h) == orc.step(h)).all()
        gpu.expand_hospital(), orc.expand_hospital()
        assert (gpu.geometry() == orc.geometry()).all()
        for h in range(49, 73):
            assert (gpu.step(h) == orc.step(h)).all()
        gpu.unlock_city(), orc.unlock_city()
        for h in range(73, 97):
            assert (gpu.step(h) == orc.step(h)).all()
        assert_state_equal(gpu, orc, "end")


def test_run_hours_graph_path_equals_single_steps():
    gc, _ = small_cfg()
    with Engine(gc, seed=6) as a, Engine(gc, seed=6) as b:
        rows = a.run_hours(1, 24 * 6 + 5)  # aligned days go through the CUDA graph, the tail does not
        for i, hour in enumerate(range(1, 24 * 6 + 6)):
            assert (rows[i] == b.step(hour)).all(), f"hour {hour}"
        sa, sb = a.get_state(), b.get_state()
        for f in STATE_FIELDS:
            assert (sa[f] == sb[f]).all()
        # unaligned start
        rows2 = a.run_hours(24 * 6 + 6, 60)
        for i, hour in enumerate(range(24 * 6 + 6, 24 * 6 + 66)):
            assert (rows2[i] == b.step(hour)).all(), f"hour {hour}"


def test_queued_hours_equal_simulate_hours():
    """epi_enqueue_hours / epi_collect_hours (the no-wait path of the multi-region day) against epi_simulate_hours: same
    Counts rows, same intervention events, same final state -- segments of odd lengths cut at epi_next_decision_hour."""
    kw = dict(n_agents=10000, grid_size=250, hours=400, exposed=100, lockdown=(60, 0.1), hospital=20, vaccinate=((100, 0.2), (131, 0.1)))
    gc = make_config(**kw)
    with Engine(gc, seed=8) as a, Engine(gc, seed=8) as b:
        want, _ = a.simulate_hours(1, 330)
        got, hour, lengths = [], 1, [5, 1, 9, 24, 3, 17, 2, 30]
        k = 0
        while hour <= 330:
            seg_end = min(330, b.next_decision_hour(hour), hour + lengths[k % len(lengths)] - 1)
            b.enqueue_hours(hour, seg_end - hour + 1)
            if seg_end != b.next_decision_hour(hour) and seg_end < 330 and k % 3 == 0:  # two segments behind one wait
                nxt = min(330, b.next_decision_hour(seg_end + 1), seg_end + 4)
                b.enqueue_hours(seg_end + 1, nxt - seg_end)
                seg_end = nxt
            rows = b.collect_hours()
            assert rows[:, 0].tolist() == list(range(hour, seg_end + 1))
            got.append(rows)
            hour, k = seg_end + 1, k + 1
        got = np.concatenate(got)
        assert got.shape == want.shape and (got == want).all(), f"first differing hour {np.nonzero((got != want).any(axis=1))[0][:3] + 1}"
        assert (a.intervention_events() == b.intervention_events()).all() and len(a.intervention_events()) >= 3
        sa, sb = a.get_state(), b.get_state()
        for f in STATE_FIELDS:
            assert (sa[f] == sb[f]).all()
        with pytest.raises(EpiError, match="queued"):  # a plain run may not overtake queued hours
            b.enqueue_hours(331, 2)
            b.run_hours(340, 1)
        b.collect_hours()


def test_whole_run_with_interventions_matches_oracle(tmp_path):
    kw = dict(n_agents=10000, grid_size=250, hours=1080, exposed=100, lockdown=(60, 0.1), hospital=20, vaccinate=((100, 0.2), (300, 0.1)))
    gc, oc = make_config(**kw), O.make_config(**kw)
    rows_g, secs = run_standalone(gc, seed=21, output_dir=str(tmp_path))
    rows_o, events_o, _ = O.oracle_run(oc, seed=21, mode="keyed", threads=4)
    assert rows_g.shape == rows_o.shape
    assert (rows_g == rows_o).all()
    import glob, json, csv
    (csv_path,) = glob.glob(str(tmp_path / "output" / "simulation_0_*[0-9].csv"))
    with open(csv_path) as f:
        rd = list(csv.reader(f))
    assert rd[0] == ["hour", "susceptible", "exposed", "infected", "hospitalized", "recovered", "deceased"]
    assert (np.array(rd[1:], dtype=np.uint32) == rows_o).all()
    (js_path,) = glob.glob(str(tmp_path / "output" / "simulation_0_*_interventions.json"))
    ev = json.load(open(js_path))
    names = {0: "lockdown", 1: "vaccination", 2: "build_new_hospital"}
    assert [(e["hour"], e["intervention"]) for e in ev] == [(int(h), names[int(k)]) for h, k, s in events_o]
    assert len(ev) >= 3
    for e, (h, k, s) in zip(ev, events_o):
        if k == 0:
            assert e["data"] == {"status": "locked_down" if s else "lockdown_revoked"}
        else:
            assert e["data"] == {}


def test_default_json_run_matches_oracle():
    """BASELINE config #1: engine/config/default.json of the reference, verbatim values."""
    gc = make_config(10000, 250, 1080, exposed=1, lockdown=(100, 0.1))
    rows_g, _ = run_standalone(gc, seed=3)
    rows_o, _, _ = O.oracle_run(O.default_json_config(), seed=3, mode="keyed", threads=4)
    assert rows_g.shape == rows_o.shape and (rows_g == rows_o).all()
