"""Oracle-compared parity at BASELINE.json's own sizes (configs #2, #3 and one region of #5), the claim-stamp epoch of long
chunks, and the HospitalStaff 14-day rotation at state level.

At these sizes the claim words carry 20-25 id bits (7-12 stamp bits), the claim array is 60-500 MB, coordinates need
12-14 bits and the agent arrays are far larger than the L2 -- none of which the 4 000-agent tests exercise.  The oracle runs
its KEYED mode on every host core (phase A in parallel, phase B in id order: the same convention as the kernels), so whole
states are bit-identical.
"""
import os

import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200.engine import Engine, make_config, STATE_FIELDS

pytestmark = pytest.mark.gpu

THREADS = max(1, os.cpu_count() or 1)


def assert_state_equal(gpu_state, orc_state, ctx, n=None):
    for f in STATE_FIELDS:
        a, b = gpu_state[f][:n], orc_state[f][:n]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, f"{ctx}: field {f} differs for {bad.size} agents, first {bad[:5]}: gpu {a[bad[:5]]} oracle {b[bad[:5]]}"


def run_both(gpu, orc, first, n):
    """hours [first, first + n): the GPU in one epi_run_hours chunk (graph replay for aligned days), the oracle hour by hour"""
    rows = gpu.run_hours(first, n)
    for k in range(n):
        want = orc.step(first + k)
        assert (rows[k] == want).all(), f"hour {first + k}: gpu {rows[k]} oracle {want}"
    return rows


def test_config2_1m_agents_72h_full_state():
    """BASELINE config #2: single region, 1 M agents, G = 2500, no interventions (id_bits 20)."""
    kw = dict(n_agents=1_000_000, grid_size=2500, hours=1080, exposed=1000)
    with Engine(make_config(**kw), seed=2) as gpu:
        orc = O.OracleEngine(O.make_config(**kw), seed=2, threads=THREADS)
        assert (gpu.geometry() == orc.geometry()).all()
        assert_state_equal(gpu.get_state(), orc.get_state(), "config #2 init")
        for day in range(3):
            rows = run_both(gpu, orc, 1 + 24 * day, 24)
            assert_state_equal(gpu.get_state(), orc.get_state(), f"config #2 end of day {day}")
        assert rows[-1, 3] > 0  # somebody became infectious within 72 h (exposed_duration 48)


def test_config3_10m_agents_state_and_interventions():
    """BASELINE config #3: 10 M agents, G = 7910 (id_bits 24, 250 MB of claim words, 13-bit coordinates), the three
    interventions' sweeps inside the compared window: vaccination at hour 30, lock_city at 48, the hospital rectangle
    switch at 48.  Full state after 24 h and at the end, Counts every hour for 72 h."""
    kw = dict(n_agents=10_000_000, grid_size=7910, hours=1080, exposed=10_000, lockdown=(100_000, 0.1), hospital=10_000, vaccinate=((30, 0.2),))
    with Engine(make_config(**kw), seed=1) as gpu:
        orc = O.OracleEngine(O.make_config(**kw), seed=1, threads=THREADS)
        assert_state_equal(gpu.get_state(), orc.get_state(), "config #3 init")
        run_both(gpu, orc, 1, 24)
        assert_state_equal(gpu.get_state(), orc.get_state(), "config #3 hour 24")
        run_both(gpu, orc, 25, 6)
        gpu.vaccinate(0.2, 30), orc.vaccinate(0.2, 30)  # allocation_map.rs:381-387 at the configured hour
        run_both(gpu, orc, 31, 18)
        gpu.lock_city(), orc.lock_city()  # allocation_map.rs:349-356
        gpu.expand_hospital(), orc.expand_hospital()  # grid.rs:233-238
        assert (gpu.geometry() == orc.geometry()).all()
        rows = run_both(gpu, orc, 49, 24)
        s = gpu.get_state()
        assert_state_equal(s, orc.get_state(), "config #3 hour 72")
        assert int(((s["st"] >> 8) & 1).sum()) > 1_500_000 and int(((s["st"] >> 11) & 1).sum()) > 8_000_000  # vaccinated, isolated
        assert int(rows[:, 1:].sum(axis=1).min()) == kw["n_agents"]


def test_config5_region_20m_agents_movement_hours():
    """One region of BASELINE config #5: 20 M agents, G = 11 180 (id_bits 25, 14-bit coordinates, 500 MB of claim words):
    full state after the movement hours 7, 8, 16 and 17 of the first day."""
    kw = dict(n_agents=20_000_000, grid_size=11_180, hours=2160, exposed=20_000)
    with Engine(make_config(**kw), seed=3) as gpu:
        orc = O.OracleEngine(O.make_config(**kw), seed=3, threads=THREADS)
        hour = 1
        for stop in (7, 8, 16, 17):
            run_both(gpu, orc, hour, stop - hour + 1)
            hour = stop + 1
            assert_state_equal(gpu.get_state(), orc.get_state(), f"config #5 region, hour {stop}")


def test_long_chunk_crosses_the_claim_stamp_limit():
    """epi_run_hours with more hours than the claim words have stamps (VERDICT r01 weak #2 / ADVICE medium): 2^25 + agent
    slots leave 6 stamp bits = 63 hours per epoch; a 200-hour chunk must be split (claim array re-zeroed in between) and
    equal 200 single steps of an engine without spare slots, row for row and in the final state."""
    kw = dict(n_agents=30_000, grid_size=440, hours=400, exposed=300, mild=30, severe=30)
    n = kw["n_agents"]
    with Engine(make_config(**kw), seed=12, extra_capacity=(1 << 25) + 5 - n) as wide, Engine(make_config(**kw), seed=12) as ref:
        assert wide.capacity == (1 << 25) + 5 and wide.population == n
        rows = wide.run_hours(1, 200)
        assert wide.epoch_resets >= 4  # 200 hours / 63 stamps
        for k in range(200):
            want = ref.step(1 + k)
            assert (rows[k] == want).all(), f"hour {1 + k}: {rows[k]} vs {want}"
        a, b = wide.get_state(), ref.get_state()
        assert_state_equal(a, b, "after 200 hours", n=n)
        moved = (a["cell_x"][:n] != b["cell_x"][:n]).sum()
        assert moved == 0
        # the queued path (epi_enqueue_hours) splits as well
        resets = wide.epoch_resets
        for first in range(201, 401, 24):
            wide.enqueue_hours(first, 24)
        got = wide.collect_hours()
        assert wide.epoch_resets > resets
        for k in range(len(got)):
            assert (got[k] == ref.step(201 + k)).all(), f"queued hour {201 + k}"
        assert_state_equal(wide.get_state(), ref.get_state(), "after the queued hours", n=n)


# ---- HospitalStaff rotation (citizen/mod.rs:292-307) ---------------------------------------------------------------------
WS_SHIFT, WS_STAFF, WS_NORMAL = 13, 2, 0
ST_WQ = 1 << 12


def test_hospital_staff_rotation_state_level():
    """`hour - work_start_at == 336` (work_quarantined := true, no infection dynamics) and `== 672` (go home, work_start_at :=
    hour + 336) on crafted states: 600 workers become HospitalStaff whose work_start_at makes both triggers fire at every
    hour of day over the next 50 hours; full state compared after every hour."""
    kw = dict(n_agents=6000, grid_size=150, hours=2000, exposed=200, asym=30, mild=30, severe=30)
    with Engine(make_config(**kw), seed=17) as gpu:
        orc = O.OracleEngine(O.make_config(**kw), seed=17)
        for h in range(1, 31):
            assert (gpu.step(h) == orc.step(h)).all()
        s = gpu.get_state()
        ws = (s["st"] >> WS_SHIFT) & 3
        normal = np.nonzero(ws == WS_NORMAL)[0][:600]
        assert normal.size == 600
        base = 1016  # the crafted state continues after hour 1016 = 8:00 (the engines keep no clock of their own): the 14 movement hours
        #              9..22 follow without an 8:00 in between, which would restart work_start_at (citizen/mod.rs:308-316)
        st = s["st"].copy()
        st[normal] = (st[normal] & ~np.uint32(3 << WS_SHIFT)) | np.uint32(WS_STAFF << WS_SHIFT)
        wsa = s["wsa"].copy()
        k = np.arange(600)
        # trigger hour = base + 1 + k % 14; a third each: since == 336, since == 672, since in {335, 671} one hour earlier
        trigger = base + 1 + k % 14
        since = np.where(k % 3 == 0, 336, np.where(k % 3 == 1, 672, 0))
        wsa[normal] = np.where(k % 3 == 2, trigger - np.where(k % 2 == 0, 335, 671), trigger - since).astype(np.uint32)
        st[normal[k % 5 == 0]] |= np.uint32(ST_WQ)  # some already work_quarantined
        s["st"], s["wsa"] = st, wsa
        gpu.set_state(s), orc.set_state(s)
        assert_state_equal(gpu.get_state(), orc.get_state(), "crafted state")
        fired_wq = fired_home = 0
        for hour in range(base + 1, base + 41):
            before = gpu.get_state()
            cg, co = gpu.step(hour), orc.step(hour)
            assert (cg == co).all(), f"hour {hour}: gpu {cg} oracle {co}"
            after = gpu.get_state()
            assert_state_equal(after, orc.get_state(), f"hour {hour}")
            d = hour - before["wsa"][normal].astype(np.int64)
            if not 7 <= hour % 24 <= 22:  # perform_movements runs at the movement hours only (citizen/mod.rs:227-255)
                d[:] = -1
            fired_wq += int((d == 336).sum())
            fired_home += int((d == 672).sum())
            went = normal[d == 672]
            assert (after["wsa"][went] == hour + 336).all()
            assert ((after["st"][normal[d == 336]] & ST_WQ) != 0).all()
        assert fired_wq >= 200 and fired_home >= 200, (fired_wq, fired_home)
