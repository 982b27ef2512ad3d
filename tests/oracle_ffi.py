"""ctypes binding of the CPU oracle (oracle/libepi_oracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libepi_oracle.so")


class EpiConfig(C.Structure):
    """Mirror of `epi_config` (include/epi.h) == `orc_config` (oracle/epi_oracle_engine.hpp)."""

    _fields_ = [
        ("number_of_agents", C.c_uint32),
        ("public_transport_percentage", C.c_double),
        ("working_percentage", C.c_double),
        ("regular_transmission_start_day", C.c_uint32),
        ("high_transmission_start_day", C.c_uint32),
        ("last_day", C.c_uint32),
        ("asymptomatic_last_day", C.c_uint32),
        ("mild_infected_last_day", C.c_uint32),
        ("regular_transmission_rate", C.c_double),
        ("high_transmission_rate", C.c_double),
        ("death_rate", C.c_double),
        ("percentage_asymptomatic_population", C.c_double),
        ("percentage_severe_infected_population", C.c_double),
        ("exposed_duration", C.c_uint32),
        ("pre_symptomatic_duration", C.c_uint32),
        ("grid_size", C.c_uint32),
        ("hospital_beds_percentage", C.c_double),
        ("hours", C.c_uint32),
        ("infected_mild_asymptomatic", C.c_uint32),
        ("infected_mild_symptomatic", C.c_uint32),
        ("infected_severe", C.c_uint32),
        ("exposed", C.c_uint32),
        ("has_lockdown", C.c_int32),
        ("lockdown_at_number_of_infections", C.c_uint32),
        ("essential_workers_population", C.c_double),
        ("has_build_new_hospital", C.c_int32),
        ("spread_rate_threshold", C.c_uint32),
        ("n_vaccinations", C.c_int32),
        ("vaccinate_at_hour", C.c_uint32 * 8),
        ("vaccinate_percent", C.c_double * 8),
        ("population_csv_file", C.c_char * 256),
    ]


# the `disease` block of engine/config/default.json (reference), used by every BASELINE config
DEFAULT_DISEASE = dict(
    regular_transmission_start_day=5, high_transmission_start_day=6, last_day=26,
    asymptomatic_last_day=9, mild_infected_last_day=12,
    regular_transmission_rate=0.25, high_transmission_rate=0.25, death_rate=0.035,
    percentage_asymptomatic_population=0.3, percentage_severe_infected_population=0.3,
    exposed_duration=48, pre_symptomatic_duration=48,
)


def make_config(n_agents=10000, grid_size=250, hours=1080, exposed=1, asym=0, mild=0, severe=0,
                pt=0.2, working=0.7, beds=0.003, lockdown=None, hospital=None, vaccinate=(), population_csv=None, **disease):
    c = EpiConfig()
    c.number_of_agents = n_agents
    c.public_transport_percentage = pt
    c.working_percentage = working
    d = dict(DEFAULT_DISEASE)
    d.update(disease)
    for k, v in d.items():
        setattr(c, k, v)
    c.grid_size = grid_size
    c.hospital_beds_percentage = beds
    c.hours = hours
    c.exposed, c.infected_mild_asymptomatic, c.infected_mild_symptomatic, c.infected_severe = exposed, asym, mild, severe
    if lockdown is not None:
        c.has_lockdown = 1
        c.lockdown_at_number_of_infections, c.essential_workers_population = lockdown
    if hospital is not None:
        c.has_build_new_hospital = 1
        c.spread_rate_threshold = hospital
    c.n_vaccinations = len(vaccinate)
    for i, (h, p) in enumerate(vaccinate):
        c.vaccinate_at_hour[i] = h
        c.vaccinate_percent[i] = p
    if population_csv:
        c.population_csv_file = str(population_csv).encode()
    return c


def default_json_config():
    """engine/config/default.json of the reference, verbatim values."""
    return make_config(10000, 250, 1080, exposed=1, lockdown=(100, 0.1))


def build_oracle(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB_PATH)
        for f in os.listdir(ORACLE_DIR) if f.endswith((".hpp", ".cpp"))
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "libepi_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(EpiConfig), C.c_uint64, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_population.restype = C.c_uint32
        L.orc_population.argtypes = [C.c_void_p]
        L.orc_set_shuffle_phase_b.argtypes = [C.c_void_p, C.c_int]
        L.orc_counts_at_start.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_step_with_draws.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_lock_city.argtypes = [C.c_void_p]
        L.orc_unlock_city.argtypes = [C.c_void_p]
        L.orc_vaccinate.argtypes = [C.c_void_p, C.c_double, C.c_uint32]
        L.orc_expand_hospital.argtypes = [C.c_void_p]
        L.orc_geometry.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_set_state.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 7
        L.orc_run.restype = C.c_long
        L.orc_run.argtypes = [C.POINTER(EpiConfig), C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_long,
                              C.c_void_p, C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_double)]
        L.orc_time_hours.restype = C.c_double
        L.orc_time_hours.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_kat_draw.restype = C.c_uint64
        L.orc_kat_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_kat_draw32.restype = C.c_uint32
        L.orc_kat_draw32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_kat_bernoulli_threshold.restype = C.c_uint64
        L.orc_kat_bernoulli_threshold.argtypes = [C.c_double]
        L.orc_kat_number_of_cells.restype = C.c_uint32
        L.orc_kat_area_iter.restype = C.c_long
        L.orc_kat_area_iter.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_long]
        L.orc_kat_area_factory.restype = C.c_long
        L.orc_kat_area_factory.argtypes = [C.c_int] * 4 + [C.c_uint32, C.c_void_p, C.c_long]
        L.orc_kat_resize_hospital.argtypes = [C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.orc_kat_transmission_rate.restype = C.c_double
        L.orc_kat_transmission_rate.argtypes = [C.POINTER(EpiConfig), C.c_uint32]
        L.orc_kat_is_to_be_hospitalized.argtypes = [C.POINTER(EpiConfig), C.c_uint32]
        L.orc_kat_goto_hospital.argtypes = [C.c_uint32, C.c_void_p, C.c_int] + [C.c_int] * 10 + [C.c_uint64, C.c_void_p]
        L.orc_iv_create.restype = C.c_void_p
        L.orc_iv_create.argtypes = [C.POINTER(EpiConfig)]
        L.orc_iv_destroy.argtypes = [C.c_void_p]
        L.orc_iv_op.restype = C.c_long
        L.orc_iv_op.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32]
        L.orc_kat_counts.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


STATE_FIELDS = ("cell_x", "cell_y", "st", "t0", "home", "work", "wsa")
STATE_DTYPES = (np.int32, np.int32, np.uint32, np.uint32, np.uint32, np.uint32, np.uint32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleEngine:
    """One region engine of the oracle.  mode: 'keyed' (Philox, id-ordered phase B) or 'stream'."""

    def __init__(self, cfg, seed=1, mode="keyed", threads=1):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.orc_create(C.byref(cfg), seed, 2 if mode == "stream" else 0, threads)
        if not self.h:
            raise RuntimeError(self.L.orc_last_error().decode())

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def population(self):
        return self.L.orc_population(self.h)

    def counts_at_start(self):
        out = np.zeros(7, np.uint32)
        self.L.orc_counts_at_start(self.h, _ptr(out))
        return out

    def step(self, hour, draws=None):
        out = np.zeros(7, np.uint32)
        if draws is None:
            rc = self.L.orc_step(self.h, hour, _ptr(out))
        else:
            draws = np.ascontiguousarray(draws, np.uint64)
            assert draws.shape == (self.population, 16)
            rc = self.L.orc_step_with_draws(self.h, hour, _ptr(draws), _ptr(out))
        if rc:
            raise RuntimeError(self.L.orc_last_error().decode())
        return out

    def lock_city(self):
        self.L.orc_lock_city(self.h)

    def unlock_city(self):
        self.L.orc_unlock_city(self.h)

    def vaccinate(self, p, hour):
        self.L.orc_vaccinate(self.h, p, hour)

    def expand_hospital(self):
        self.L.orc_expand_hospital(self.h)

    def geometry(self):
        out = np.zeros(19, np.int32)
        self.L.orc_geometry(self.h, _ptr(out))
        return out

    def get_state(self):
        n = self.population
        arrs = {f: np.zeros(n, dt) for f, dt in zip(STATE_FIELDS, STATE_DTYPES)}
        if self.L.orc_get_state(self.h, *[_ptr(arrs[f]) for f in STATE_FIELDS]):
            raise RuntimeError(self.L.orc_last_error().decode())
        return arrs

    def set_state(self, arrs):
        n = len(arrs["st"])
        a = [np.ascontiguousarray(arrs[f], dt) for f, dt in zip(STATE_FIELDS, STATE_DTYPES)]
        if self.L.orc_set_state(self.h, n, *[_ptr(x) for x in a]):
            raise RuntimeError(self.L.orc_last_error().decode())

    def time_hours(self, first_hour, n_hours):
        return self.L.orc_time_hours(self.h, first_hour, n_hours)


def oracle_run(cfg, seed=1, mode="keyed", threads=1, max_hours=0):
    """Whole standalone run.  Returns (rows[n,7], events[m,3], hour-loop seconds)."""
    L = lib()
    max_rows = int(cfg.hours)
    rows = np.zeros((max_rows, 7), np.uint32)
    events = np.zeros((64, 3), np.uint32)
    ne = C.c_long(0)
    secs = C.c_double(0)
    n = L.orc_run(C.byref(cfg), seed, 2 if mode == "stream" else 0, threads, max_hours, _ptr(rows), max_rows, _ptr(events), 64,
                  C.byref(ne), C.byref(secs))
    if n < 0:
        raise RuntimeError(L.orc_last_error().decode())
    return rows[:n].copy(), events[: ne.value].copy(), secs.value


# ---- multi-region oracle (oracle/epi_oracle_travel.hpp) -----------------------------------------------------------------
MULTI_STATE_FIELDS = STATE_FIELDS + ("reg",)
MULTI_STATE_DTYPES = STATE_DTYPES + (np.uint32,)


def _multi_lib():
    L = lib()
    if not getattr(L, "_multi_ready", False):
        L.orc_multi_create.restype = C.c_void_p
        L.orc_multi_create.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_multi_destroy.argtypes = [C.c_void_p]
        L.orc_multi_step.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_multi_capacity.restype = C.c_uint32
        L.orc_multi_capacity.argtypes = [C.c_void_p, C.c_int]
        L.orc_multi_population.restype = C.c_uint32
        L.orc_multi_population.argtypes = [C.c_void_p, C.c_int]
        L.orc_multi_max_place_rounds.restype = C.c_uint32
        L.orc_multi_max_place_rounds.argtypes = [C.c_void_p, C.c_int]
        L.orc_multi_get_state.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
        L.orc_multi_events.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_kat_percent_outgoing.restype = C.c_double
        L.orc_kat_percent_outgoing.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32]
        L.orc_kat_alloc_outgoing.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        L._multi_ready = True
    return L


class OracleMultiEngine:
    """R region engines in lock step with the traveller exchange (Epidemiology::run_multi_engine)."""

    def __init__(self, cfgs, seed, migration=None, commute=None, start_migration_hour=0, end_migration_hour=0, extra_capacity=0, threads=1):
        self.L = _multi_lib()
        self.R = len(cfgs)
        arr = (EpiConfig * self.R)(*cfgs)
        mig = np.ascontiguousarray(migration if migration is not None else np.zeros((self.R, self.R)), np.uint32)
        com = np.ascontiguousarray(commute if commute is not None else np.zeros((self.R, self.R)), np.uint32)
        self.h = self.L.orc_multi_create(C.cast(arr, C.c_void_p), self.R, seed, _ptr(mig), _ptr(com), int(migration is not None), int(commute is not None),
                                         start_migration_hour, end_migration_hour, extra_capacity, threads)
        if not self.h:
            raise RuntimeError(self.L.orc_last_error().decode())

    def close(self):
        if self.h:
            self.L.orc_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def step(self, hour):
        rows = np.zeros((self.R, 7), np.uint32)
        if self.L.orc_multi_step(self.h, hour, _ptr(rows)):
            raise RuntimeError(self.L.orc_last_error().decode())
        return rows

    def capacity(self, r):
        return self.L.orc_multi_capacity(self.h, r)

    def population(self, r):
        return self.L.orc_multi_population(self.h, r)

    def max_place_rounds(self, r):
        """most placement rounds one arrival batch of region r needed so far (select_starting_points)"""
        return self.L.orc_multi_max_place_rounds(self.h, r)

    def get_state(self, r):
        n = self.capacity(r)
        arrs = {f: np.zeros(n, dt) for f, dt in zip(MULTI_STATE_FIELDS, MULTI_STATE_DTYPES)}
        if self.L.orc_multi_get_state(self.h, r, *[_ptr(arrs[f]) for f in MULTI_STATE_FIELDS]):
            raise RuntimeError(self.L.orc_last_error().decode())
        return arrs

    def events(self, r):
        ev = np.zeros((64, 3), np.uint32)
        n = self.L.orc_multi_events(self.h, r, _ptr(ev), 64)
        return ev[:n]
