"""ctypes binding of the C ABI in include/epi.h (epirust_b200/libepirust_b200.so).

Fails loudly when the CUDA library is missing: there is no CPU fallback in the product.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# EPI_LIB: developer override to A/B-test an alternative build of the same CUDA library (tools_exp.py)
LIB_PATH = os.environ.get("EPI_LIB") or os.path.join(PKG, "libepirust_b200.so")

EPI_DRAWS_PER_AGENT = 16
EPI_N_KERNEL_KINDS = 8
KERNEL_KINDS = ("hour", "commit", "hospital_scan", "sleep", "sweep", "spare", "travel", "misc")


class EpiConfig(C.Structure):
    """`epi_config` of include/epi.h (== common::config::Config for Population::Auto)."""

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
        ("population_csv_file", C.c_char * 256),  # Population::Csv.file, b"" for Population::Auto
    ]


class EpiTravelPlan(C.Structure):
    """`epi_travel_plan` of include/epi.h (common::config::TravelPlanConfig, regions named by index)."""

    _fields_ = [("n_regions", C.c_int32), ("migration_enabled", C.c_int32), ("commute_enabled", C.c_int32), ("migration", C.c_void_p),
                ("commute", C.c_void_p), ("start_migration_hour", C.c_uint32), ("end_migration_hour", C.c_uint32)]


COMM_ID_BYTES = 128
TRAVEL_RECORD_BYTES = 32
TRAVEL_MIGRATE, TRAVEL_COMMUTE = 0, 1


class EpiCounts(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("hour", "susceptible", "exposed", "infected", "hospitalized", "recovered", "deceased")]


# every symbol include/epi.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "epi_create", "epi_create_region", "epi_create_multi", "epi_destroy", "epi_last_error", "epi_population", "epi_capacity", "epi_counts_at_start",
    "epi_set_stream", "epi_travel_pack", "epi_travel_unpack", "epi_finish_hour", "epi_get_regions",
    "epi_sync", "epi_reset", "epi_step", "epi_enqueue_hour", "epi_enqueue_hours", "epi_collect_hours", "epi_next_decision_hour", "epi_step_with_draws", "epi_run_hours", "epi_simulate_hours", "epi_intervention_events", "epi_lock_city", "epi_unlock_city", "epi_vaccinate",
    "epi_expand_hospital", "epi_get_state", "epi_set_state", "epi_build_population", "epi_population_size", "epi_geometry", "epi_get_grid", "epi_set_kernel_timing",
    "epi_get_kernel_times", "epi_get_hour_times", "epi_launch_count", "epi_device_bytes", "epi_epoch_resets", "epi_set_tiles", "epi_tile_hours", "epi_debug_trace", "epi_config_from_json", "epi_config_from_json_string",
    "epi_run_standalone", "epi_run_standalone_ex", "epi_config_citizen_state_messages", "epi_citizen_states", "epi_version", "epi_device_count",
    "epi_comm_unique_id", "epi_comm_init", "epi_comm_init_local", "epi_comm_destroy", "epi_exchange_kind", "epi_exchange", "epi_run_multi_hours",
    "epi_count_outgoing", "epi_outgoing_travels", "epi_should_terminate", "epi_multi_schedule_trace",
    "epi_configuration_read", "epi_configuration_free", "epi_configuration_regions", "epi_configuration_region_name", "epi_configuration_engine_config",
    "epi_configuration_travel_plan", "epi_configuration_arrival_capacity", "epi_run_region", "epi_write_outputs",
]

_lib = None


def _preload_bundled_nccl():
    """libepirust_b200.so links libnccl.so.2 (the traveller exchange).  PyTorch wheels ship their own, newer libnccl.so.2 and
    libtorch_cuda.so needs symbols only that one has: whichever copy a process loads first serves both.  So in a Python process
    that might import torch later (tests, bench.py) the bundled copy is loaded first when it exists; a process without PyTorch
    (the engine-app binary) uses the system library."""
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        p = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(p):
            C.CDLL(p, mode=C.RTLD_GLOBAL)
            return p
    return None


def load():
    """Load the CUDA shared library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m epirust_b200.build` (nvcc, sm_100a). "
            "epirust_b200 has no CPU fallback.")
    _preload_bundled_nccl()
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.epi_create.argtypes = [C.POINTER(EpiConfig), u64, i32, C.POINTER(vp)]
    L.epi_create_region.argtypes = [C.POINTER(EpiConfig), u64, i32, i32, C.POINTER(vp)]
    L.epi_create_multi.argtypes = [C.POINTER(EpiConfig), u64, i32, i32, C.POINTER(EpiTravelPlan), u32, C.POINTER(vp)]
    L.epi_capacity.argtypes = [vp]
    L.epi_capacity.restype = u32
    L.epi_travel_pack.argtypes = [vp, u32, i32, vp, u64, vp]
    L.epi_travel_unpack.argtypes = [vp, u32, i32, vp, u64, vp]
    L.epi_finish_hour.argtypes = [vp, u32, C.POINTER(EpiCounts)]
    L.epi_get_regions.argtypes = [vp, vp]
    L.epi_destroy.argtypes = [vp]
    L.epi_destroy.restype = None
    L.epi_last_error.argtypes = [vp]
    L.epi_last_error.restype = C.c_char_p
    L.epi_population.argtypes = [vp]
    L.epi_population.restype = u32
    L.epi_counts_at_start.argtypes = [vp, C.POINTER(EpiCounts)]
    L.epi_set_stream.argtypes = [vp, vp]
    L.epi_sync.argtypes = [vp]
    L.epi_reset.argtypes = [vp]
    L.epi_step.argtypes = [vp, u32, C.POINTER(EpiCounts)]
    L.epi_enqueue_hour.argtypes = [vp, u32]
    L.epi_enqueue_hours.argtypes = [vp, u32, u32]
    L.epi_collect_hours.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.epi_next_decision_hour.argtypes = [vp, u32]
    L.epi_next_decision_hour.restype = u32
    L.epi_step_with_draws.argtypes = [vp, u32, vp, C.POINTER(EpiCounts)]
    L.epi_run_hours.argtypes = [vp, u32, u32, vp]
    L.epi_simulate_hours.argtypes = [vp, u32, u32, i32, vp, C.POINTER(u32), C.POINTER(i32)]
    L.epi_intervention_events.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.epi_lock_city.argtypes = [vp]
    L.epi_unlock_city.argtypes = [vp]
    L.epi_vaccinate.argtypes = [vp, C.c_double, u32]
    L.epi_expand_hospital.argtypes = [vp]
    L.epi_get_state.argtypes = [vp] + [vp] * 7
    L.epi_set_state.argtypes = [vp, u32] + [vp] * 7
    L.epi_build_population.argtypes = [C.POINTER(EpiConfig), u64] + [vp] * 7
    L.epi_population_size.argtypes = [C.POINTER(EpiConfig), C.POINTER(u32)]
    L.epi_geometry.argtypes = [vp, vp]
    L.epi_get_grid.argtypes = [vp, vp, u64, C.POINTER(u32), C.POINTER(u32)]
    L.epi_set_kernel_timing.argtypes = [vp, i32]
    L.epi_get_kernel_times.argtypes = [vp, vp, vp]
    L.epi_launch_count.argtypes = [vp, i32]
    L.epi_launch_count.restype = u64
    L.epi_device_bytes.argtypes = [vp]
    L.epi_device_bytes.restype = u64
    L.epi_epoch_resets.argtypes = [vp]
    L.epi_epoch_resets.restype = u64
    L.epi_get_hour_times.argtypes = [vp, vp, vp]
    L.epi_set_tiles.argtypes = [vp, i32]
    L.epi_tile_hours.argtypes = [vp]
    L.epi_tile_hours.restype = u64
    L.epi_debug_trace.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.epi_config_from_json.argtypes = [C.c_char_p, C.POINTER(EpiConfig)]
    L.epi_config_from_json_string.argtypes = [C.c_char_p, C.POINTER(EpiConfig)]
    L.epi_run_standalone.argtypes = [C.POINTER(EpiConfig), u64, i32, C.c_char_p, C.c_char_p, vp, u32, C.POINTER(u32), C.POINTER(C.c_double)]
    L.epi_run_standalone_ex.argtypes = [C.POINTER(EpiConfig), u64, i32, C.c_char_p, C.c_char_p, i32, vp, u32, C.POINTER(u32), C.POINTER(C.c_double)]
    L.epi_config_citizen_state_messages.argtypes = [C.c_char_p, C.POINTER(i32)]
    L.epi_citizen_states.argtypes = [vp, vp, vp, vp, vp, u32, C.POINTER(u32)]
    L.epi_version.restype = C.c_char_p
    L.epi_comm_unique_id.argtypes = [vp]
    L.epi_comm_init.argtypes = [vp, i32, i32, vp]
    L.epi_comm_init_local.argtypes = [vp, i32]
    L.epi_comm_destroy.argtypes = [vp]
    L.epi_exchange_kind.argtypes = [vp, u32]
    L.epi_exchange.argtypes = [vp, u32, i32]
    L.epi_run_multi_hours.argtypes = [vp, i32, u32, u32, i32, vp, C.POINTER(u32)]
    L.epi_count_outgoing.argtypes = [vp, i32]
    L.epi_outgoing_travels.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.epi_should_terminate.argtypes = [vp, i32]
    L.epi_multi_schedule_trace.argtypes = [C.POINTER(EpiTravelPlan), vp, i32, u32, u32, u32, C.c_char_p, u64]
    L.epi_configuration_read.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.epi_configuration_free.argtypes = [vp]
    L.epi_configuration_free.restype = None
    L.epi_configuration_regions.argtypes = [vp]
    L.epi_configuration_region_name.argtypes = [vp, i32]
    L.epi_configuration_region_name.restype = C.c_char_p
    L.epi_configuration_engine_config.argtypes = [vp, i32, C.POINTER(EpiConfig)]
    L.epi_configuration_travel_plan.argtypes = [vp, i32, C.POINTER(EpiTravelPlan), vp, vp]
    L.epi_configuration_arrival_capacity.argtypes = [vp, i32]
    L.epi_configuration_arrival_capacity.restype = u32
    L.epi_run_region.argtypes = [vp, i32, i32, vp, u64, i32, C.c_char_p, i32, vp, u32, C.POINTER(u32), C.POINTER(C.c_double)]
    L.epi_write_outputs.argtypes = [C.c_char_p, C.c_char_p, vp, u32, vp, u32, vp, u32, vp, C.c_char_p, u64]
    _lib = L
    return L
