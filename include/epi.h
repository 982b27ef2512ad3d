/*
 * epi.h -- C ABI of the B200-native EpiRust per-hour agent step.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference (thoughtworks/epirust) has no
 * FFI: its hour loop calls Rust methods directly.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference root) so a maintainer can swap the call site for an `extern "C"`
 * binding (see INTEGRATION.md for the Rust `extern` block).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types.
 *   - every call returns 0 on success, non-zero on failure; epi_last_error() gives the message.  Nothing
 *     panics or throws across the boundary (the reference panics: allocation_map.rs:128, :222, :254 ...).
 *   - one epi_engine == one region == one GPU + one CUDA stream.  A handle is NOT thread-safe (one caller
 *     thread per handle, like the reference: epidemiology_simulation.rs:223-257).
 *   - agents and the occupancy grid live in HBM for the whole run.  Only the 7 x u32 Counts row per hour
 *     (and traveller buffers on exchange hours) cross PCIe.
 *   - there is NO CPU fallback: every call fails with EPI_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef EPI_H
#define EPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPI_OK 0
#define EPI_ERR_ARG 1
#define EPI_ERR_CUDA 2
#define EPI_ERR_CONFIG 3
#define EPI_ERR_IO 4
#define EPI_ERR_STATE 5
#define EPI_ERR_NCCL 6

#define EPI_MAX_VACCINATIONS 8
#define EPI_PATH_MAX 256
#define EPI_DRAWS_PER_AGENT 16 /* u64 draw slots per agent-hour, see DESIGN.md "Draw slots" */

/* common::config::Config for Population::Auto (common/src/config/mod.rs:44-58, population.rs:36-43,
 * common/src/disease/mod.rs:26-45, geography_parameters.rs:24-28, starting_infections.rs:22-28,
 * intervention_config.rs:23-48).  Produced from the simulation-config JSON by epi_config_from_json(). */
typedef struct epi_config {
    uint32_t number_of_agents;
    double public_transport_percentage;
    double working_percentage;
    uint32_t regular_transmission_start_day, high_transmission_start_day, last_day;
    uint32_t asymptomatic_last_day, mild_infected_last_day; /* parsed, ignored like the reference (constants.rs:48-50) */
    double regular_transmission_rate, high_transmission_rate, death_rate;
    double percentage_asymptomatic_population, percentage_severe_infected_population;
    uint32_t exposed_duration, pre_symptomatic_duration;
    uint32_t grid_size;
    double hospital_beds_percentage;
    uint32_t hours;
    uint32_t infected_mild_asymptomatic, infected_mild_symptomatic, infected_severe, exposed;
    int32_t has_lockdown;
    uint32_t lockdown_at_number_of_infections;
    double essential_workers_population;
    int32_t has_build_new_hospital;
    uint32_t spread_rate_threshold;
    int32_t n_vaccinations;
    uint32_t vaccinate_at_hour[EPI_MAX_VACCINATIONS];
    double vaccinate_percent[EPI_MAX_VACCINATIONS];
    /* Population::Csv { file, cols } (common/src/config/population.rs:30-34): path of the population file, or "" for
     * Population::Auto.  When set, epi_create* reads it like Grid::read_population (engine/src/geography/grid.rs:194-231):
     * one PopulationRecord per line (columns ind, age, working, pub_transport; citizen/population_record.rs:23-31), record c
     * gets house c % H and office c % O; number_of_agents and the two Auto percentages are ignored.  `cols` and
     * `disease_overrides` are parsed and ignored, as in the reference. */
    char population_csv_file[EPI_PATH_MAX];
} epi_config;

/* engine::models::events::Counts (engine/src/models/events/counts.rs:25-34): one epicurve CSV row. */
typedef struct epi_counts {
    uint32_t hour, susceptible, exposed, infected, hospitalized, recovered, deceased;
} epi_counts;

/* one InterventionReport (engine/src/listeners/intervention_reporter.rs:28-33).  kind: 0 lockdown, 1 vaccination,
 * 2 build_new_hospital (InterventionType::name, lockdown.rs:104-114, vaccination.rs:58-64, hospital.rs:78-84);
 * status: lockdown 1 = "locked_down", 0 = "lockdown_revoked"; otherwise 0 (json_data is {}). */
typedef struct epi_intervention_event {
    uint32_t hour;
    int32_t kind;
    int32_t status;
} epi_intervention_event;

/* common::config::TravelPlanConfig (common/src/config/travel_plan_config.rs:22-41) with regions named by their index in
 * `regions` (== rank == GPU).  Matrices are n_regions x n_regions, row-major [from][to]. */
typedef struct epi_travel_plan {
    int32_t n_regions;
    int32_t migration_enabled, commute_enabled;
    const uint32_t* migration;
    const uint32_t* commute;
    uint32_t start_migration_hour, end_migration_hour;
} epi_travel_plan;
#define EPI_TRAVEL_RECORD_BYTES 32 /* one traveller on the wire (Commuter / Migrator, engine/src/travel) */
#define EPI_TRAVEL_MIGRATE 0
#define EPI_TRAVEL_COMMUTE 1

typedef struct epi_engine epi_engine;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */

/* Replaces Epidemiology::new (engine/src/epidemiology_simulation.rs:75-135): define_geography, Auto population
 * factory, resize_hospital, CitizenLocationMap::new, init_interventions' essential-worker draw.  `seed` is the
 * Philox key (the reference is unseeded: common/src/utils/random_wrapper.rs:23-35).  `region` is this engine's
 * index in the travel plan (0 for standalone). */
int epi_create(const epi_config* cfg, uint64_t seed, int device, epi_engine** out);
int epi_create_region(const epi_config* cfg, uint64_t seed, int device, int region, epi_engine** out);
/* One region of a multi-region run (Epidemiology::new with a travel plan + EngineApp::start_with_mpi's per-rank setup,
 * engine-app/src/main.rs:131-166): additionally applies citizen_factory::update_commuters (citizen_factory.rs:90-110),
 * builds the house / office occupancy heaps (grid.rs:125-155, 262-277) and reserves `extra_capacity` empty agent slots
 * for arrivals.  The engine's Philox key is `seed` (callers pass a different seed per region). */
int epi_create_multi(const epi_config* cfg, uint64_t seed, int device, int region, const epi_travel_plan* plan, uint32_t extra_capacity,
                     epi_engine** out);
void epi_destroy(epi_engine* e);
/* message of the last failed call on this handle (NULL handle: last failed epi_create / global call) */
const char* epi_last_error(const epi_engine* e);
/* CitizenLocationMap::current_population (allocation_map.rs:389-391) */
uint32_t epi_population(const epi_engine* e);
/* number of agent slots (== population for a standalone engine) */
uint32_t epi_capacity(const epi_engine* e);
/* counts_at_start (engine/src/utils/util.rs:45-51) */
int epi_counts_at_start(const epi_engine* e, epi_counts* out);
/* use the caller's CUDA stream (cudaStream_t as void*) for all subsequent work; NULL = the engine's own */
int epi_set_stream(epi_engine* e, void* cuda_stream);
int epi_sync(epi_engine* e);
/* restore the state epi_create produced (bench: re-run from hour 1 without paying init again) */
int epi_reset(epi_engine* e);

/* ---- the hot path ----------------------------------------------------------------------------------------- */

/* Replaces CitizenLocationMap::simulate (engine/src/allocation_map.rs:67-129) for one simulated hour:
 * Citizen::perform_operation for every agent against the start-of-hour map, lowest-agent-id conflict
 * resolution, swap, and Counts::update_counts.  out->hour = hour. */
int epi_step(epi_engine* e, uint32_t hour, epi_counts* out);
/* epi_step without waiting for the hour to finish and without its Counts row: the multi-region loop enqueues the exchange hour,
 * packs the leavers behind it on the same stream and reads the row from epi_finish_hour. */
int epi_enqueue_hour(epi_engine* e, uint32_t hour);
/* A multi-region day without intermediate host waits.  epi_enqueue_hours queues the hours [first_hour, first_hour + n_hours)
 * -- consecutive with what is already queued, none of them an exchange hour -- and returns at once (replaying a CUDA graph
 * per (hour of day, n_hours)); their Counts rows stay on the device.  epi_collect_hours waits for everything queued
 * (including an exchange hour queued by epi_enqueue_hour and its pack / unpack), returns the rows of the non-exchange hours
 * in order and runs process_interventions on each, like epi_simulate_hours; the exchange hour's own row then comes from
 * epi_finish_hour.  A queued segment must end at epi_next_decision_hour() at the latest: the next hour >= `hour` whose
 * Counts the host has to see before the following hour may run (start of day: lockdown.rs:55, hospital.rs:55,70; a
 * configured vaccination hour: vaccination.rs:52; the unlock hour: lockdown.rs:69-73). */
int epi_enqueue_hours(epi_engine* e, uint32_t first_hour, uint32_t n_hours);
int epi_collect_hours(epi_engine* e, epi_counts* rows_out, uint32_t max_rows, uint32_t* n_rows);
uint32_t epi_next_decision_hour(const epi_engine* e, uint32_t hour);
/* Same hour with every random draw injected: draws[agent * EPI_DRAWS_PER_AGENT + slot] (host memory).
 * This is the bit-exact sub-step test entry (no reference equivalent; the reference cannot inject draws). */
int epi_step_with_draws(epi_engine* e, uint32_t hour, const uint64_t* draws, epi_counts* out);
/* n_hours consecutive epi_step calls without host synchronisation in between (CUDA-graph replay for whole
 * days).  rows_out[n_hours].  Replaces the body of the loop at epidemiology_simulation.rs:223-246 between
 * intervention decisions. */
int epi_run_hours(epi_engine* e, uint32_t first_hour, uint32_t n_hours, epi_counts* rows_out);

/* The body of the hour loop of Epidemiology::run_single_engine (epidemiology_simulation.rs:223-257) for hours
 * first_hour .. first_hour+n_hours-1: counts.increment_hour, simulate, listeners.counts_updated,
 * process_interventions (allocation_map.rs:306-337: vaccinate / lock / unlock / build hospital, decided on the host from
 * the Counts row exactly like interventions/ *.rs, applied by the sweep kernels), and, when `stop_rule` is non-zero,
 * Epidemiology::stop_simulation's Standalone arm (:564-575).  The device runs ahead to the next hour at which a host
 * decision can change device state (start of day, a configured vaccination hour, the unlock hour), so Counts cross PCIe
 * once per simulated day.  *n_rows = rows written (< n_hours only if the stop rule fired; *stopped = 1 then). */
int epi_simulate_hours(epi_engine* e, uint32_t first_hour, uint32_t n_hours, int stop_rule, epi_counts* rows_out, uint32_t* n_rows,
                       int* stopped);
/* InterventionReporter's list since epi_create / epi_reset; out may be NULL to query *n only */
int epi_intervention_events(const epi_engine* e, epi_intervention_event* out, uint32_t max_events, uint32_t* n);

/* ---- multi-region: the traveller exchange (engine/src/epidemiology_simulation.rs:391-503) -------------------------------
 * Per exchange hour (h % 24 == 0 migrators inside the migration window, h % 24 in {7, 17} commuters) the caller runs
 *   epi_step(hour) -> epi_travel_pack -> all-to-all of the segments (the caller's transport; epi_exchange below is this
 *   library's own, and the one the hour loop uses) -> epi_travel_unpack -> epi_finish_hour.
 * Replaces Transport::send_* / receive_* (engine/src/transport/mod.rs:34-42, mpi_transport.rs:78-215) around
 * remove_* / assimilate_* (allocation_map.rs:165-277).  Records never leave device memory and all of the reference's
 * sequential bookkeeping (allotment of migrators to regions, free agent slots, house / office occupancy heaps) runs on the
 * device; the host only sees the per-region record counts.
 *
 * Buffer layout (send and receive alike, DEVICE memory): n_regions segments of stride_records records of
 * EPI_TRAVEL_RECORD_BYTES each.  Record 0 of a segment is a header whose first 32-bit word is the number of records that
 * follow (<= stride_records - 1).  In the send buffer segment d is addressed to region d; after an all-to-all with equal
 * splits segment s of the receive buffer holds what region s sent here. */
/* Selects the leaving agents (Citizen::is_commuter / can_migrate + gen_bool(percent_outgoing), citizen/mod.rs:456-495),
 * allots migrators to regions (EngineMigrationPlan::alloc_outgoing_to_regions, engine_migration_plan.rs:51-77), writes
 * their records into the destination's segment of send_buf and removes them from the region.  counts_out[n_regions]
 * (HOST) = records per destination.  Every segment header is written, also when nobody travels.
 * counts_out == NULL defers: the call returns with the kernels in flight on the engine's stream (the collective can be
 * queued behind them at once) and errors surface in the next epi_finish_hour. */
int epi_travel_pack(epi_engine* e, uint32_t hour, int kind, void* send_buf, uint64_t stride_records, uint32_t* counts_out);
/* Installs the arrivals of recv_buf in order of source region.  assimilate_migrators / assimilate_commuters
 * (allocation_map.rs:214-277).  counts_in[n_regions] (HOST) receives the records per source region; counts_in == NULL
 * defers like epi_travel_pack (epi_finish_hour settles, including further placement rounds if some arrival is still
 * without a cell). */
int epi_travel_unpack(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, uint64_t stride_records, uint32_t* counts_in);
/* The tail of the multi-engine hour (epidemiology_simulation.rs:492-503): Counts after the travel adjustments,
 * process_interventions, and stop_simulation's MultiEngine arm (:564-571, records lockdown.zero_infection_hour). */
int epi_finish_hour(epi_engine* e, uint32_t hour, epi_counts* out);
/* home region | work region << 8 per agent slot (0 for empty slots): Area.location_id of home_location / work_location */
int epi_get_regions(epi_engine* e, uint32_t* reg);

/* ---- multi-region: the Transport and the hour loop behind the ABI ---------------------------------------------------------
 * Replaces the reference's `Transport` trait (engine/src/transport/mod.rs:34-42) and its MPI implementation
 * (MpiTransport::send_commuters / send_migrators / receive_commuters / receive_migrators, transport/mpi_transport.rs:78-215:
 * bincode + snappy over MPI point-to-point) by an all-to-allv of packed 32-byte records between the GPUs, one rank per region,
 * everything on the engine's stream with no host wait inside.  Two data planes, chosen at epi_comm_init by all ranks together:
 *   peer memory (default): every rank maps the others' receive areas and flag words (CUDA IPC); an exchange is ONE cooperative
 *     kernel that removes the leavers, stores each destination's records straight into that rank's receive area over NVLink /
 *     NVSwitch, raises its flag there, waits for the peers' flags and installs the arrivals (csrc/travel.cu, k_travel_exchange);
 *   NCCL (when a peer cannot be mapped, or EPI_NO_PEER=1): leave kernel -> grouped ncclSend / ncclRecv -> arrive kernel.
 * Each (source, destination) pair ships only the records it has (peer memory) or the records the travel plan can produce for it
 * (NCCL: the commute matrix entry; the migration matrix entry + 8 sigma), not a padded segment.
 *
 * epi_comm_unique_id: ncclGetUniqueId.  One rank calls it and hands the EPI_COMM_ID_BYTES bytes to the others by any means
 *   (a file, an environment variable, MPI_Bcast: the reference's ranks meet through mpirun, engine-app/src/main.rs:131-147).
 * epi_comm_init: ncclCommInitRank on the engine's device; rank must equal the engine's region index, n_ranks the number of
 *   regions of its travel plan (MpiTransport::new, mpi_transport.rs:44-52).  Allocates the send / receive segments.
 * epi_comm_init_local: every region of the plan lives in this process (tests; several regions on one GPU, or one process
 *   driving several GPUs): the "collective" is a set of device-to-device copies ordered by CUDA events.
 * Failures of NCCL itself return EPI_ERR_NCCL with ncclGetErrorString in epi_last_error. */
#define EPI_COMM_ID_BYTES 128
int epi_comm_unique_id(void* id_out);
int epi_comm_init(epi_engine* e, int n_ranks, int rank, const void* unique_id);
int epi_comm_init_local(epi_engine* const* engines, int n_engines);
int epi_comm_destroy(epi_engine* e);
/* EPI_TRAVEL_MIGRATE / EPI_TRAVEL_COMMUTE when `hour` is an exchange hour of this engine's travel plan, -1 otherwise
 * (MpiTransport::receive_tick, mpi_transport.rs:60-76; migrators only inside the migration window, citizen/mod.rs:460-462) */
int epi_exchange_kind(const epi_engine* e, uint32_t hour);
/* One traveller exchange on an engine with a communicator from epi_comm_init: leave -> records to the peers -> arrive (one
 * cooperative kernel over peer memory, or pack -> ncclSend / ncclRecv -> unpack), all deferred: counts, population and errors
 * stay on the device and surface at the next epi_collect_hours / epi_finish_hour.  Every rank of the communicator must call it for the same (hour, kind).  The body of
 * epidemiology_simulation.rs:407-488. */
int epi_exchange(epi_engine* e, uint32_t hour, int kind);
/* Epidemiology::run_multi_engine's hour loop (epidemiology_simulation.rs:331-537) for hours first_hour .. first_hour +
 * n_hours - 1 of the n_local engines this process hosts (1 with epi_comm_init, all regions with epi_comm_init_local):
 * the plain hours between two exchanges are queued as CUDA graphs, the exchange hour's kernels, pack, collective and unpack
 * follow on the same stream, and the host waits once per exchange / decision hour.  rows_out[n_local][n_hours].
 * terminate_when_clear != 0 adds the orchestrator's global termination rule (orchestrator/src/ticks.rs:35-89, 175-180; Kafka
 * mode -- MPI mode always runs to `hours`): at every tick hour (hour 1 and the hours of day 0 / 7 / 17 whose exchange kind is
 * enabled) the regions' exposed + infected + hospitalized are summed over all ranks; when the sum is zero the run stops
 * before the next tick hour.  *n_rows = hours executed (the same on every rank). */
int epi_run_multi_hours(epi_engine* const* engines, int n_local, uint32_t first_hour, uint32_t n_hours, int terminate_when_clear, epi_counts* rows_out,
                        uint32_t* n_rows);
/* TravelCounter (engine/src/listeners/travel_counter.rs:27-92): one CountsByRegion per destination the plan sends migrators to,
 * at every hour % 24 == 0 of a migration-enabled run.  epi_count_outgoing(e, 1) turns the listener on (the used part of the send
 * segments then also goes to the host behind the pack kernels); epi_outgoing_travels reads what accumulated since epi_create /
 * epi_reset (out may be NULL to query *n). */
typedef struct epi_outgoing_travel {
    uint32_t hr, destination, susceptible, exposed, infected, recovered; /* destination: region index in the travel plan */
} epi_outgoing_travel;
int epi_count_outgoing(epi_engine* e, int on);
int epi_outgoing_travels(const epi_engine* e, epi_outgoing_travel* out, uint32_t max_rows, uint32_t* n);
/* TickAcks::should_terminate (orchestrator/src/ticks.rs:175-180) on the Counts the regions acknowledged for one tick: 1 when
 * no region has exposed, infected or hospitalized agents.  Host only. */
int epi_should_terminate(const epi_counts* acks, int n_acks);
/* Host only (no GPU): the order in which epi_run_multi_hours queues work for one region whose decision hours are start of day
 * plus `vaccinate_hours` / `unlock_hour` (0 = none), written to `out` as text, one call per line ("hours F N", "exchange_hour
 * H", "exchange H KIND", "collect", "finish H").  Test hook of the scheduling logic. */
int epi_multi_schedule_trace(const epi_travel_plan* plan, const uint32_t* vaccinate_hours, int n_vaccinate, uint32_t unlock_hour, uint32_t first_hour,
                             uint32_t n_hours, char* out, uint64_t out_bytes);

/* ---- interventions: the O(N) sweeps; the decisions stay with the host (interventions/ *.rs) ---------------- */
/* CitizenLocationMap::lock_city (allocation_map.rs:349-356) */
int epi_lock_city(epi_engine* e);
/* CitizenLocationMap::unlock_city (allocation_map.rs:358-365) */
int epi_unlock_city(epi_engine* e);
/* CitizenLocationMap::vaccinate (allocation_map.rs:381-387); `hour` keys the draws */
int epi_vaccinate(epi_engine* e, double vaccination_percentage, uint32_t hour);
/* Grid::increase_hospital_size (engine/src/geography/grid.rs:233-238) */
int epi_expand_hospital(epi_engine* e);

/* ---- test harness: crafted states in, states out ------------------------------------------------------------ */
/* Arrays of length epi_population(), in agent-id order.  st is the packed state word (DESIGN.md "Agent state
 * word"), t0 = at_hour of Exposed / Pre, home/work = house/office index, wsa = HospitalStaff.work_start_at. */
int epi_get_state(epi_engine* e, int32_t* cell_x, int32_t* cell_y, uint32_t* st, uint32_t* t0, uint32_t* home, uint32_t* work,
                  uint32_t* wsa);
int epi_set_state(epi_engine* e, uint32_t n, const int32_t* cell_x, const int32_t* cell_y, const uint32_t* st, const uint32_t* t0,
                  const uint32_t* home, const uint32_t* work, const uint32_t* wsa);
/* Host only (no GPU needed): the Auto population factory -- Grid::generate_population + citizen_factory +
 * set_starting_infections + the essential-worker draw of init_interventions (engine/src/geography/grid.rs:83-155,
 * citizen/citizen_factory.rs:31-134, epidemiology_simulation.rs:178-192; Grid::read_population, grid.rs:194-231, when
 * cfg->population_csv_file is set) -- as the arrays epi_create uploads, in the layout of epi_get_state.  Arrays of
 * epi_population_size() entries. */
/* Host only: the number of agents `cfg` describes -- number_of_agents, or the number of records of the population CSV. */
int epi_population_size(const epi_config* cfg, uint32_t* n);
int epi_build_population(const epi_config* cfg, uint64_t seed, int32_t* cell_x, int32_t* cell_y, uint32_t* st, uint32_t* t0,
                         uint32_t* home, uint32_t* work, uint32_t* wsa);
/* Listener::citizen_state_updated (engine/src/listeners/listener.rs:33; EventsKafkaProducer, listeners/events_kafka_producer.rs:90-100;
 * CitizenState, models/events/citizen_state.rs:26-62): what the reference publishes per agent and hour on the
 * `citizen_states_updated` topic when Config.enable_citizen_state_messages is set -- the state letter ('s', 'e', 'i', 'r', 'd':
 * CitizenState::state_str) and the location of every live agent after the last simulated hour, in slot order; slot_out (may be
 * NULL) receives the slot, this engine's stand-in for Citizen.id.  Up to `capacity` agents are written; *n_out = live agents. */
int epi_citizen_states(epi_engine* e, char* state_out, int32_t* x_out, int32_t* y_out, uint32_t* slot_out, uint32_t capacity, uint32_t* n_out);
/* out[0..3] housing sx,sy,ex,ey; [4..7] transport; [8..11] work; [12..15] hospital (current); [16] houses; [17] offices;
 * [18] grid_size.  (geography/mod.rs:33-70, grid.rs:240-261) */
int epi_geometry(const epi_engine* e, int32_t* out19);
/* raw occupancy grid bytes (pitch * rows) for invariants tests; *pitch, *rows receive the dimensions */
int epi_get_grid(epi_engine* e, uint8_t* out, uint64_t capacity, uint32_t* pitch, uint32_t* rows);

/* ---- measurement ------------------------------------------------------------------------------------------------ */
#define EPI_N_KERNEL_KINDS 8
/* kinds: 0 hour kernel (propose + transition + counts + claim), 1 commit (lowest-id claim resolution), 2 hospital scan,
 * 3 sleep/area-reset, 4 intervention sweeps, 5 (unused), 6 traveller pack/unpack, 7 misc.  When timing is on every launch is bracketed by CUDA events on the engine's
 * stream (this serialises nothing but adds event overhead; never on during the bench's headline timing). */
int epi_set_kernel_timing(epi_engine* e, int on);
int epi_get_kernel_times(epi_engine* e, double* ms_total, uint64_t* launches);
/* the same per hour of day: [h * 2 + 0] the hour's agent kernels (one k_hour launch, or generic segment + tile kernel), [h * 2 + 1] its
 * commit pass; 48 entries each */
int epi_get_hour_times(epi_engine* e, double* ms_total, uint64_t* launches);
/* number of kernel launches (graph kernel nodes included) since creation / last reset of the counter */
uint64_t epi_launch_count(const epi_engine* e, int reset);
/* device bytes held by the engine */
uint64_t epi_device_bytes(const epi_engine* e);
/* How often the claim array was zeroed because the hour stamps of the claim words (32 - id_bits bits, id_bits =
 * ceil(log2(agent slots))) were used up: 255 hours at 10 M agents.  epi_run_hours / epi_enqueue_hours split their work at
 * that limit.  (No reference equivalent: the reference clears its `upcoming` map every hour, allocation_map.rs:131-134.) */
uint64_t epi_epoch_resets(const epi_engine* e);
/* Optional tile kernels for the plain movement hours (h % 24 in 9..11, 13..15, 18..22) of a standalone engine (csrc/tiles.cu: one
 * warp per run of adjacent offices / houses, the run's grid bytes staged into shared memory by TMA bulk copies, lowest-id claims
 * settled on chip, no claim[] traffic and no commit pass for the members).  Same results bit for bit; off by default because the
 * id-order kernels are faster on B200.  epi_set_tiles switches them on / off (environment EPI_TILES=1 = on at
 * creation), epi_tile_hours counts the hours that took them.  (No reference equivalent.) */
int epi_set_tiles(epi_engine* e, int on);
uint64_t epi_tile_hours(const epi_engine* e);
/* Debugging: with EPI_TRACE=1 in the environment at creation the kernels leave nanosecond stamps (pairs: (tag << 56) | time, hour; tag 1 = a
 * commit pass starts, 2 / 3 = the exchange's leave kernel starts / ends, 4 / 5 / 6 = its arrive kernel starts / has every peer's travellers / ends).
 * Returns and clears them. */
int epi_debug_trace(epi_engine* e, uint64_t* out, uint64_t max_words, uint64_t* n_words);

/* ---- host driver: the engine-app equivalent ----------------------------------------------------------------- */
/* Parse the reference's simulation-config JSON (common::config::Config::read, common/src/config/mod.rs:124-128). */
int epi_config_from_json(const char* json_path, epi_config* out);
int epi_config_from_json_string(const char* json_text, epi_config* out);
/* EngineApp::start_standalone (engine/src/engine_app.rs:89-106) + Epidemiology::run_single_engine
 * (epidemiology_simulation.rs:211-274): runs the whole simulation and writes
 * <output_dir>/output/simulation_<engine_id>_<UTC>.csv and ..._interventions.json like CsvListener
 * (listeners/csv_service.rs:44-71) and InterventionReporter (listeners/intervention_reporter.rs:28-63).
 * rows_out (may be NULL) receives up to max_rows rows; *n_rows the number of hours executed; *loop_seconds the
 * hour-loop wall time (what the reference logs as Iterations/sec, epidemiology_simulation.rs:270-272). */
int epi_run_standalone(const epi_config* cfg, uint64_t seed, int device, const char* output_dir, const char* engine_id,
                       epi_counts* rows_out, uint32_t max_rows, uint32_t* n_rows, double* loop_seconds);
/* The same with Config.enable_citizen_state_messages (common/src/config/mod.rs:54-55, default false): when
 * citizen_state_messages != 0 the run proceeds hour by hour and appends to <output_dir>/output/simulation_<engine_id>_<UTC>
 * _citizen_states.jsonl one line per simulated hour, the JSON the reference's EventsKafkaProducer sends to the
 * `citizen_states_updated` topic (serde of CitizenStatesAtHr: {"hr":H,"citizen_states":[{"citizen_id":"..","state":"s",
 * "location":{"x":X,"y":Y}}, ..]}; citizen_id is a UUID-shaped rendering of the agent slot, the reference's is a random Uuid),
 * and {"simulation_ended": true} as the last line (events_kafka_producer.rs:77-88).  There is no broker here: the file is the topic. */
int epi_run_standalone_ex(const epi_config* cfg, uint64_t seed, int device, const char* output_dir, const char* engine_id, int citizen_state_messages,
                          epi_counts* rows_out, uint32_t max_rows, uint32_t* n_rows, double* loop_seconds);
/* Host only: Config.enable_citizen_state_messages of a simulation-config JSON file (absent = 0, serde default) */
int epi_config_citizen_state_messages(const char* json_path, int* on);

/* ---- host driver: multi-region runs (engine-app -m mpi) ----------------------------------------------------------------- */
/* common::config::Configuration (common/src/config/configuration.rs:28-57): {engine_configs: [{engine_id, config}],
 * travel_plan: {regions, migration: {enabled, matrix, start_migration_hour, end_migration_hour}, commute: {enabled, matrix}}}.
 * epi_configuration_read = Configuration::read + TravelPlanConfig::validate_regions (travel_plan_config.rs:60-62) +
 * Configuration::validate (configuration.rs:59-117); what the reference panics on comes back as EPI_ERR_CONFIG.  Regions are
 * named by their index in travel_plan.regions == rank (MpiTransport::new, transport/mpi_transport.rs:44-52). */
typedef struct epi_configuration epi_configuration;
int epi_configuration_read(const char* json_path, epi_configuration** out);
void epi_configuration_free(epi_configuration* c);
int epi_configuration_regions(const epi_configuration* c);
const char* epi_configuration_region_name(const epi_configuration* c, int region);
/* the Config of the engine whose engine_id names `region` */
int epi_configuration_engine_config(const epi_configuration* c, int region, epi_config* out);
/* the travel plan restricted to its first n_regions regions; migration_out / commute_out: n_regions x n_regions words each that
 * the returned plan points to */
int epi_configuration_travel_plan(const epi_configuration* c, int n_regions, epi_travel_plan* out, uint32_t* migration_out, uint32_t* commute_out);
/* agent slots epi_run_region reserves for the arrivals of `region` (every commuter of a day + the migrators of the window) */
uint32_t epi_configuration_arrival_capacity(const epi_configuration* c, int region);
/* One rank of `engine-app -m mpi` (engine-app/src/main.rs:131-166: rank r takes engine_configs[r], EngineApp::start_with_mpi,
 * Epidemiology::run_multi_engine): creates region `region` of the first n_ranks regions on `device` with Philox key seed +
 * region, joins the communicator `unique_id` names (epi_comm_unique_id), runs hours 1 .. config.hours - 1 and writes
 * <output_dir>/output/simulation_<engine_id>_<UTC>.csv, ..._interventions.json, ..._outgoing_travels.csv (output_dir may be
 * NULL).  terminate_when_clear: see epi_run_multi_hours. */
int epi_run_region(const epi_configuration* c, int region, int n_ranks, const void* unique_id, uint64_t seed, int device, const char* output_dir,
                   int terminate_when_clear, epi_counts* rows_out, uint32_t max_rows, uint32_t* n_rows, double* loop_seconds);

/* Host only: what the listeners write at simulation_ended -- <output_dir>/output/simulation_<engine_id>_<UTC>.csv (CsvListener,
 * listeners/csv_service.rs:44-71), ..._interventions.json (InterventionReporter, intervention_reporter.rs:28-63) and, when
 * `travels` is not NULL, ..._outgoing_travels.csv (TravelCounter, travel_counter.rs:69-79; region_names[destination] names the
 * regions).  base_out (may be NULL) receives the path without its suffix. */
int epi_write_outputs(const char* output_dir, const char* engine_id, const epi_counts* rows, uint32_t n_rows, const epi_intervention_event* events,
                      uint32_t n_events, const epi_outgoing_travel* travels, uint32_t n_travels, const char* const* region_names, char* base_out,
                      uint64_t base_bytes);
/* number of usable CUDA devices (0 when there is none; the library has no CPU fallback) */
int epi_device_count(void);
const char* epi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EPI_H */
