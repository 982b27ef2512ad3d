"""Python mirror of the region engine over the C ABI (tests, bench and the torchrun launcher use this).

The class and method names follow the reference's own vocabulary (CitizenLocationMap::simulate, lock_city,
vaccinate ...; engine/src/allocation_map.rs) so the parity tests read like the reference's tests.
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import EpiConfig, EpiCounts

STATE_FIELDS = ("cell_x", "cell_y", "st", "t0", "home", "work", "wsa")
STATE_DTYPES = (np.int32, np.int32, np.uint32, np.uint32, np.uint32, np.uint32, np.uint32)
COUNT_FIELDS = ("hour", "susceptible", "exposed", "infected", "hospitalized", "recovered", "deceased")

# the `disease` block of the reference's engine/config/default.json
DEFAULT_DISEASE = dict(
    regular_transmission_start_day=5, high_transmission_start_day=6, last_day=26,
    asymptomatic_last_day=9, mild_infected_last_day=12,
    regular_transmission_rate=0.25, high_transmission_rate=0.25, death_rate=0.035,
    percentage_asymptomatic_population=0.3, percentage_severe_infected_population=0.3,
    exposed_duration=48, pre_symptomatic_duration=48,
)


def make_config(n_agents=10000, grid_size=250, hours=1080, exposed=1, asym=0, mild=0, severe=0,
                pt=0.2, working=0.7, beds=0.003, lockdown=None, hospital=None, vaccinate=(), population_csv=None, **disease):
    """Build an EpiConfig; defaults are engine/config/default.json without its Lockdown intervention.  population_csv:
    path of a population file (Population::Csv) -- n_agents, pt and working are then ignored."""
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


def config_from_json(path):
    L = _ffi.load()
    c = EpiConfig()
    if L.epi_config_from_json(str(path).encode(), C.byref(c)):
        raise ValueError(L.epi_last_error(None).decode())
    return c


def config_from_json_string(text):
    L = _ffi.load()
    c = EpiConfig()
    if L.epi_config_from_json_string(text.encode(), C.byref(c)):
        raise ValueError(L.epi_last_error(None).decode())
    return c


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def counts_to_array(c):
    return np.array([getattr(c, f) for f in COUNT_FIELDS], np.uint32)


class EpiError(RuntimeError):
    pass


class Engine:
    """One region engine resident on one GPU (`epi_engine`)."""

    def __init__(self, cfg, seed=1, device=0, region=0, plan=None, extra_capacity=0):
        """plan: None (standalone) or dict(n_regions, migration=RxR or None, commute=RxR or None, start_migration_hour, end_migration_hour)."""
        self.L = _ffi.load()
        self.cfg = cfg
        h = C.c_void_p()
        if plan is None and extra_capacity:  # a standalone engine with empty agent slots (tests: a wide id field in the claim words)
            rc = self.L.epi_create_multi(C.byref(cfg), seed, device, region, None, extra_capacity, C.byref(h))
        elif plan is None:
            rc = self.L.epi_create_region(C.byref(cfg), seed, device, region, C.byref(h))
        else:
            R = int(plan["n_regions"])
            self._mig = np.ascontiguousarray(plan.get("migration") if plan.get("migration") is not None else np.zeros((R, R)), np.uint32)
            self._com = np.ascontiguousarray(plan.get("commute") if plan.get("commute") is not None else np.zeros((R, R)), np.uint32)
            assert self._mig.shape == (R, R) and self._com.shape == (R, R)
            tp = _ffi.EpiTravelPlan(R, int(plan.get("migration") is not None), int(plan.get("commute") is not None), self._mig.ctypes.data,
                                    self._com.ctypes.data, int(plan.get("start_migration_hour", 0)), int(plan.get("end_migration_hour", 0)))
            self.n_regions = R
            rc = self.L.epi_create_multi(C.byref(cfg), seed, device, region, C.byref(tp), extra_capacity, C.byref(h))
        if rc:
            raise EpiError(f"epi_create failed ({rc}): {self.L.epi_last_error(None).decode()}")
        self.h = h
        self.stream_ptr = 0  # the engine's own stream

    def _check(self, rc):
        if rc:
            raise EpiError(f"error {rc}: {self.L.epi_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.epi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def population(self):
        return self.L.epi_population(self.h)

    @property
    def capacity(self):
        return self.L.epi_capacity(self.h)

    # ---- traveller exchange (multi-region engines) ----
    def travel_pack(self, hour, kind, send_ptr, stride_records, want_counts=True):
        """send_ptr: device pointer (int) of n_regions segments of stride_records records (segment d = header + the records for
        region d).  Returns the per-destination record counts (host, numpy uint32[n_regions]); with want_counts=False the call is
        deferred (kernels in flight, settled by finish_hour) and returns None."""
        if not want_counts:
            self._check(self.L.epi_travel_pack(self.h, hour, kind, C.c_void_p(send_ptr), stride_records, None))
            return None
        counts = np.zeros(self.n_regions, np.uint32)
        self._check(self.L.epi_travel_pack(self.h, hour, kind, C.c_void_p(send_ptr), stride_records, _ptr(counts)))
        return counts

    def travel_unpack(self, hour, kind, recv_ptr, stride_records, want_counts=True):
        """recv_ptr: device pointer of n_regions segments (segment s = header + the records region s sent).  Returns counts per source
        (None when deferred)."""
        if not want_counts:
            self._check(self.L.epi_travel_unpack(self.h, hour, kind, C.c_void_p(recv_ptr), stride_records, None))
            return None
        counts_in = np.zeros(self.n_regions, np.uint32)
        self._check(self.L.epi_travel_unpack(self.h, hour, kind, C.c_void_p(recv_ptr), stride_records, _ptr(counts_in)))
        return counts_in

    def finish_hour(self, hour):
        c = EpiCounts()
        self._check(self.L.epi_finish_hour(self.h, hour, C.byref(c)))
        return counts_to_array(c)

    def get_regions(self):
        reg = np.zeros(self.capacity, np.uint32)
        self._check(self.L.epi_get_regions(self.h, _ptr(reg)))
        return reg

    def counts_at_start(self):
        c = EpiCounts()
        self._check(self.L.epi_counts_at_start(self.h, C.byref(c)))
        return counts_to_array(c)

    def set_stream(self, cuda_stream_ptr):
        self._check(self.L.epi_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))
        self.stream_ptr = int(cuda_stream_ptr or 0)

    def enqueue_hour(self, hour):
        """CitizenLocationMap::simulate for one hour, asynchronously (no Counts row; see epi_finish_hour)."""
        self._check(self.L.epi_enqueue_hour(self.h, hour))

    def enqueue_hours(self, first_hour, n_hours):
        """Queue hours [first_hour, first_hour + n_hours) without waiting (epi_enqueue_hours); rows come from collect_hours()."""
        self._check(self.L.epi_enqueue_hours(self.h, first_hour, n_hours))

    def collect_hours(self, max_rows=2400):
        """Wait for the queued hours; their Counts rows [n, 7] in order (an exchange hour's row comes from finish_hour)."""
        rows = np.zeros((max_rows, 7), np.uint32)
        n = C.c_uint32(0)
        self._check(self.L.epi_collect_hours(self.h, _ptr(rows), max_rows, C.byref(n)))
        return rows[: n.value]

    def next_decision_hour(self, hour):
        return int(self.L.epi_next_decision_hour(self.h, hour))

    def sync(self):
        self._check(self.L.epi_sync(self.h))

    def reset(self):
        self._check(self.L.epi_reset(self.h))

    # CitizenLocationMap::simulate for one hour
    def step(self, hour, draws=None):
        c = EpiCounts()
        if draws is None:
            self._check(self.L.epi_step(self.h, hour, C.byref(c)))
        else:
            draws = np.ascontiguousarray(draws, np.uint64)
            assert draws.shape == (self.capacity, _ffi.EPI_DRAWS_PER_AGENT)
            self._check(self.L.epi_step_with_draws(self.h, hour, _ptr(draws), C.byref(c)))
        return counts_to_array(c)

    def run_hours(self, first_hour, n_hours, out=None):
        rows = out if out is not None else np.zeros((n_hours, 7), np.uint32)
        self._check(self.L.epi_run_hours(self.h, first_hour, n_hours, _ptr(rows)))
        return rows

    # the hour-loop body of Epidemiology::run_single_engine: simulate + process_interventions (+ stop rule)
    def simulate_hours(self, first_hour, n_hours, stop_rule=False, out=None):
        rows = out if out is not None else np.zeros((n_hours, 7), np.uint32)
        n, stopped = C.c_uint32(0), C.c_int(0)
        self._check(self.L.epi_simulate_hours(self.h, first_hour, n_hours, int(stop_rule), _ptr(rows), C.byref(n), C.byref(stopped)))
        return rows[: n.value], bool(stopped.value)

    def intervention_events(self):
        n = C.c_uint32(0)
        self._check(self.L.epi_intervention_events(self.h, None, 0, C.byref(n)))
        ev = np.zeros((n.value, 3), np.int32)
        if n.value:
            self._check(self.L.epi_intervention_events(self.h, _ptr(ev), n.value, C.byref(n)))
        return ev

    def lock_city(self):
        self._check(self.L.epi_lock_city(self.h))

    def unlock_city(self):
        self._check(self.L.epi_unlock_city(self.h))

    def vaccinate(self, p, hour):
        self._check(self.L.epi_vaccinate(self.h, p, hour))

    def expand_hospital(self):
        self._check(self.L.epi_expand_hospital(self.h))

    def get_state(self):
        n = self.capacity
        arrs = {f: np.zeros(n, dt) for f, dt in zip(STATE_FIELDS, STATE_DTYPES)}
        self._check(self.L.epi_get_state(self.h, *[_ptr(arrs[f]) for f in STATE_FIELDS]))
        return arrs

    def citizen_states(self):
        """Listener::citizen_state_updated: (state letters as bytes 's','e','i','r','d', x, y, slot) of every live agent."""
        n = self.capacity
        state, x, y, slot = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint32)
        live = C.c_uint32()
        self._check(self.L.epi_citizen_states(self.h, _ptr(state), _ptr(x), _ptr(y), _ptr(slot), n, C.byref(live)))
        k = live.value
        return state[:k], x[:k], y[:k], slot[:k]

    def set_state(self, arrs):
        a = [np.ascontiguousarray(arrs[f], dt) for f, dt in zip(STATE_FIELDS, STATE_DTYPES)]
        self._check(self.L.epi_set_state(self.h, len(a[0]), *[_ptr(x) for x in a]))

    def geometry(self):
        out = np.zeros(19, np.int32)
        self._check(self.L.epi_geometry(self.h, _ptr(out)))
        return out

    def get_grid(self):
        pitch, rows = C.c_uint32(), C.c_uint32()
        self._check(self.L.epi_get_grid(self.h, None, 0, C.byref(pitch), C.byref(rows)))
        g = np.zeros((rows.value, pitch.value), np.uint8)
        self._check(self.L.epi_get_grid(self.h, _ptr(g), g.size, C.byref(pitch), C.byref(rows)))
        return g

    def set_kernel_timing(self, on):
        self._check(self.L.epi_set_kernel_timing(self.h, int(on)))

    def kernel_times(self):
        ms = np.zeros(_ffi.EPI_N_KERNEL_KINDS, np.float64)
        n = np.zeros(_ffi.EPI_N_KERNEL_KINDS, np.uint64)
        self._check(self.L.epi_get_kernel_times(self.h, _ptr(ms), _ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(_ffi.KERNEL_KINDS)}

    def hour_times(self):
        """per hour of day: (ms of the hour's agent kernels, launches, ms of its commit pass, launches) -- needs set_kernel_timing(True)"""
        ms = np.zeros(48, np.float64)
        n = np.zeros(48, np.uint64)
        self._check(self.L.epi_get_hour_times(self.h, _ptr(ms), _ptr(n)))
        return {h: (float(ms[2 * h]), int(n[2 * h]), float(ms[2 * h + 1]), int(n[2 * h + 1])) for h in range(24) if n[2 * h] or n[2 * h + 1]}

    def launch_count(self, reset=False):
        return int(self.L.epi_launch_count(self.h, int(reset)))

    # ---- multi-region: the Transport behind the C ABI (include/epi.h "multi-region" block) ----
    def comm_init(self, n_ranks, rank, unique_id):
        """Join the NCCL communicator `unique_id` names (bytes from comm_unique_id()); rank == this engine's region."""
        assert len(unique_id) == _ffi.COMM_ID_BYTES
        buf = C.create_string_buffer(bytes(unique_id), _ffi.COMM_ID_BYTES)
        self._check(self.L.epi_comm_init(self.h, n_ranks, rank, buf))

    def comm_destroy(self):
        self._check(self.L.epi_comm_destroy(self.h))

    def exchange_kind(self, hour):
        """TRAVEL_MIGRATE / TRAVEL_COMMUTE when `hour` is an exchange hour of the travel plan, else None."""
        k = self.L.epi_exchange_kind(self.h, hour)
        return None if k < 0 else k

    def exchange(self, hour, kind):
        """pack -> NCCL all-to-allv -> unpack on the engine's stream (deferred; finish_hour settles)."""
        self._check(self.L.epi_exchange(self.h, hour, kind))

    def count_outgoing(self, on=True):
        self._check(self.L.epi_count_outgoing(self.h, int(on)))

    def outgoing_travels(self):
        """TravelCounter rows [n, 6]: hr, destination region index, susceptible, exposed, infected, recovered."""
        n = C.c_uint32(0)
        self._check(self.L.epi_outgoing_travels(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 6), np.uint32)
        if n.value:
            self._check(self.L.epi_outgoing_travels(self.h, _ptr(out), n.value, C.byref(n)))
        return out

    def debug_trace(self):
        """EPI_TRACE=1: [(tag, nanoseconds, hour)] stamps left by the kernels since the last call"""
        buf = np.zeros(1 << 16, np.uint64)
        n = C.c_uint64(0)
        self._check(self.L.epi_debug_trace(self.h, _ptr(buf), len(buf), C.byref(n)))
        w = buf[: n.value].reshape(-1, 2)
        return [(int(a >> np.uint64(56)), int(a & np.uint64((1 << 56) - 1)), int(h)) for a, h in w]

    def set_tiles(self, on):
        """tile kernels of the plain movement hours on / off (same results either way)"""
        self._check(self.L.epi_set_tiles(self.h, int(on)))

    @property
    def tile_hours(self):
        return int(self.L.epi_tile_hours(self.h))

    @property
    def device_bytes(self):
        return int(self.L.epi_device_bytes(self.h))

    @property
    def epoch_resets(self):
        return int(self.L.epi_epoch_resets(self.h))


def population_size(cfg):
    """Number of agents `cfg` describes: number_of_agents, or the number of records of its population CSV (host only)."""
    L = _ffi.load()
    n = C.c_uint32(0)
    if L.epi_population_size(C.byref(cfg), C.byref(n)):
        raise EpiError(L.epi_last_error(None).decode(errors="replace"))
    return n.value


def build_population(cfg, seed=1):
    """The population factory on the host (epi_build_population; no GPU needed): dict of arrays like Engine.get_state()."""
    L = _ffi.load()
    arrs = {f: np.zeros(population_size(cfg), dt) for f, dt in zip(STATE_FIELDS, STATE_DTYPES)}
    if L.epi_build_population(C.byref(cfg), seed, *[_ptr(arrs[f]) for f in STATE_FIELDS]):
        raise EpiError(L.epi_last_error(None).decode(errors="replace"))
    return arrs


def run_standalone(cfg, seed=1, device=0, output_dir=None, engine_id="0"):
    """EngineApp::start_standalone: whole run with interventions; returns (rows[n,7], hour-loop seconds)."""
    L = _ffi.load()
    rows = np.zeros((max(int(cfg.hours), 1), 7), np.uint32)
    n = C.c_uint32(0)
    secs = C.c_double(0.0)
    rc = L.epi_run_standalone(C.byref(cfg), seed, device, output_dir.encode() if output_dir else None, engine_id.encode(),
                              _ptr(rows), rows.shape[0], C.byref(n), C.byref(secs))
    if rc:
        raise EpiError(f"epi_run_standalone failed ({rc}): {L.epi_last_error(None).decode()}")
    return rows[: n.value].copy(), secs.value


# ---- multi-region host driver over the C ABI ------------------------------------------------------------------------------
def comm_unique_id():
    """ncclGetUniqueId (epi_comm_unique_id): 128 bytes one rank creates and hands to the others."""
    L = _ffi.load()
    buf = C.create_string_buffer(_ffi.COMM_ID_BYTES)
    if L.epi_comm_unique_id(buf):
        raise EpiError(L.epi_last_error(None).decode())
    return buf.raw


def device_count():
    return int(_ffi.load().epi_device_count())


def travel_plan_struct(plan):
    """dict(n_regions, migration=RxR or None, commute=RxR or None, start_migration_hour, end_migration_hour) -> (EpiTravelPlan, keep-alive arrays)"""
    R = int(plan["n_regions"])
    mig = np.ascontiguousarray(plan.get("migration") if plan.get("migration") is not None else np.zeros((R, R)), np.uint32)
    com = np.ascontiguousarray(plan.get("commute") if plan.get("commute") is not None else np.zeros((R, R)), np.uint32)
    tp = _ffi.EpiTravelPlan(R, int(plan.get("migration") is not None), int(plan.get("commute") is not None), mig.ctypes.data, com.ctypes.data,
                            int(plan.get("start_migration_hour", 0)), int(plan.get("end_migration_hour", 0)))
    return tp, (mig, com)


def run_multi_hours(engines, first_hour, n_hours, terminate_when_clear=False):
    """epi_run_multi_hours: Epidemiology::run_multi_engine's hour loop for the engines this process hosts (one with an NCCL
    communicator, or every region of a local one).  Returns rows[n_local, n_rows, 7]; n_rows < n_hours when the termination rule fired."""
    L = _ffi.load()
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    rows = np.zeros((len(engines), n_hours, 7), np.uint32)
    n = C.c_uint32(0)
    rc = L.epi_run_multi_hours(arr, len(engines), first_hour, n_hours, int(terminate_when_clear), _ptr(rows), C.byref(n))
    if rc:
        raise EpiError(f"error {rc}: {L.epi_last_error(engines[0].h).decode()}")
    return rows[:, : n.value]


def should_terminate(acks):
    """TickAcks::should_terminate (orchestrator/src/ticks.rs:175-180) on Counts rows [n, 7]."""
    a = np.ascontiguousarray(acks, np.uint32).reshape(-1, 7)
    return bool(_ffi.load().epi_should_terminate(_ptr(a), len(a)))


def multi_schedule_trace(plan, first_hour, n_hours, vaccinate_hours=(), unlock_hour=0):
    """The order in which epi_run_multi_hours queues work for one region (host only): list of tuples
    ("hours", first, n) / ("exchange_hour", h) / ("exchange", h, kind) / ("collect", [hours]) / ("finish", h)."""
    L = _ffi.load()
    tp, keep = travel_plan_struct(plan)
    vac = np.ascontiguousarray(vaccinate_hours, np.uint32)
    out = C.create_string_buffer(1 << 20)
    if L.epi_multi_schedule_trace(C.byref(tp), _ptr(vac), len(vac), int(unlock_hour or 0), first_hour, n_hours, out, len(out)):
        raise EpiError(L.epi_last_error(None).decode())
    calls = []
    for line in out.value.decode().splitlines():
        w = line.split()
        calls.append((w[0], [int(v) for v in w[1:]]) if w[0] == "collect" else (w[0], *[int(v) for v in w[1:]]))
    return calls


class Configuration:
    """common::config::Configuration (configuration.rs:28-117) through epi_configuration_read: read + validate."""

    def __init__(self, path):
        self.L = _ffi.load()
        h = C.c_void_p()
        if self.L.epi_configuration_read(str(path).encode(), C.byref(h)):
            raise EpiError(self.L.epi_last_error(None).decode())
        self.h = h
        self.n_regions = self.L.epi_configuration_regions(h)
        self.regions = [self.L.epi_configuration_region_name(h, r).decode() for r in range(self.n_regions)]

    def close(self):
        if getattr(self, "h", None):
            self.L.epi_configuration_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def engine_config(self, region):
        c = EpiConfig()
        if self.L.epi_configuration_engine_config(self.h, region, C.byref(c)):
            raise EpiError(self.L.epi_last_error(None).decode())
        return c

    def travel_plan(self, n_regions=None):
        """The plan of the first n_regions regions as the dict Engine(plan=...) takes."""
        R = n_regions or self.n_regions
        tp = _ffi.EpiTravelPlan()
        mig, com = np.zeros((R, R), np.uint32), np.zeros((R, R), np.uint32)
        if self.L.epi_configuration_travel_plan(self.h, R, C.byref(tp), _ptr(mig), _ptr(com)):
            raise EpiError(self.L.epi_last_error(None).decode())
        return dict(n_regions=R, regions=self.regions[:R], migration=mig if tp.migration_enabled else None, commute=com if tp.commute_enabled else None,
                    start_migration_hour=int(tp.start_migration_hour), end_migration_hour=int(tp.end_migration_hour))

    def arrival_capacity(self, region):
        return int(self.L.epi_configuration_arrival_capacity(self.h, region))

    def run_region(self, region, n_ranks, unique_id, seed=1, device=0, output_dir=None, terminate_when_clear=False):
        """One rank of `engine-app -m mpi` (epi_run_region).  Returns (rows[n, 7], hour-loop seconds)."""
        hours = int(self.engine_config(region).hours)
        rows = np.zeros((max(hours, 1), 7), np.uint32)
        n, secs = C.c_uint32(0), C.c_double(0.0)
        buf = C.create_string_buffer(bytes(unique_id), _ffi.COMM_ID_BYTES)
        rc = self.L.epi_run_region(self.h, region, n_ranks, buf, seed, device, output_dir.encode() if output_dir else None, int(terminate_when_clear),
                                   _ptr(rows), rows.shape[0], C.byref(n), C.byref(secs))
        if rc:
            raise EpiError(f"epi_run_region failed ({rc}): {self.L.epi_last_error(None).decode()}")
        return rows[: n.value].copy(), secs.value


def write_outputs(output_dir, engine_id, rows, events, travels=None, region_names=()):
    """epi_write_outputs: the listeners' files (CSV, interventions JSON, optionally outgoing travels).  Returns the base path."""
    L = _ffi.load()
    rows = np.ascontiguousarray(rows, np.uint32).reshape(-1, 7)
    ev = np.ascontiguousarray(events, np.int32).reshape(-1, 3)
    tr = None if travels is None else np.ascontiguousarray(travels, np.uint32).reshape(-1, 6)
    names = (C.c_char_p * max(1, len(region_names)))(*[n.encode() for n in region_names])
    base = C.create_string_buffer(4096)
    rc = L.epi_write_outputs(str(output_dir).encode(), engine_id.encode(), _ptr(rows), len(rows), _ptr(ev), len(ev), None if tr is None else _ptr(tr),
                             0 if tr is None else len(tr), names, base, len(base))
    if rc:
        raise EpiError(L.epi_last_error(None).decode())
    return base.value.decode()
