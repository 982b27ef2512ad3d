// One region engine: owns the HBM-resident state of a region and a CUDA stream; implements the C ABI of include/epi.h.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/epi.h"
#include "host_model.h"
#include "layout.h"
#include "simulation.h"

namespace epi {
struct Comm;  // multi.h: the Transport of a multi-region engine
}

struct epi_engine {
    explicit epi_engine(const epi_config& c) : cfg(c), interventions(c) {}
    epi_config cfg{};
    epi::Geometry geo{};
    epi::Params P{};
    epi::DevPtrs D{};
    int device = 0;
    uint64_t seed = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // device allocations
    uint8_t* grid_alloc = nullptr;  // the cell grid with its zero padding (== D.grid)
    epi::Clock* d_clock = nullptr;
    uint32_t* d_misc = nullptr;  // [0] hosp_first, [1] collisions
    uint64_t* d_draws = nullptr;
    size_t draws_capacity = 0;
    // snapshot of the initial agent state for epi_reset (device copies)
    uint32_t *i_cell = nullptr, *i_st = nullptr, *i_t0 = nullptr, *i_home = nullptr, *i_work = nullptr, *i_wsa = nullptr;
    uint32_t* h_counts = nullptr;  // pinned staging for the counts ring
    uint64_t device_bytes = 0;
    // claim stamping
    uint32_t epoch_base = 0;
    bool claim_dirty = true;  // claim array needs zeroing before next use
    uint64_t epoch_resets = 0;  // times the claim array was zeroed and the stamp epoch restarted (test hook: epi_epoch_resets)
    // day graph (hours h%24 = 1..23,0)
    cudaGraphExec_t day_graph = nullptr;
    uint32_t day_graph_launches = 0;
    bool graphs_enabled = true;
    // hours queued by epi_enqueue_hours / epi_enqueue_hour whose Counts rows are still in the device ring (epi_collect_hours):
    // ring row k = hour pend_first + k; pend_kind[k]: 0 row repeats the previous one (sleep hour without its own k_sleep),
    // 1 row produced by the hour's kernels, 2 exchange hour (its row comes from epi_finish_hour)
    uint32_t pend_first = 0;
    std::vector<uint8_t> pend_kind;
    std::vector<uint32_t> pend_population;  // live agents when the hour was queued (allocation_map.rs:128 check at collect time)
    struct SegmentGraph {
        cudaGraphExec_t exec = nullptr;
        uint32_t launches = 0, n_sleep = 0, n_active = 0, n_scan = 0, n_tile = 0;
    };
    std::vector<std::pair<uint32_t, SegmentGraph>> segment_graphs;  // key = (first hour of day) * 32 + hours (1..24)
    // tile kernels of the plain movement hours (tiles.cu): [0] office tiles / order A (h = 9..11, 13..15), [1] house tiles / order B (h = 18..22)
    bool tiles_enabled = false;  // standalone engines unless EPI_TILES=0
    bool tiles_ready = false;    // the orders match the agents' home / work / work-status words
    epi::TileGeom tile_geom[2]{};
    epi::TilePtrs tile_ptrs[2]{};
    uint32_t tile_generic_bound[2] = {0, 0};  // launch size of the generic segment (>= its length)
    uint32_t *tile_keys_a = nullptr, *tile_keys_b = nullptr, *tile_ids = nullptr, *d_tile_misc = nullptr;
    void* tile_temp = nullptr;
    size_t tile_temp_bytes = 0;
    uint64_t tile_hours = 0;  // hours that ran on the tile kernels (test hook: epi_tile_hours)
    // measurement
    bool timing = false;
    double kernel_ms[EPI_N_KERNEL_KINDS] = {0};
    uint64_t kernel_launches[EPI_N_KERNEL_KINDS] = {0};
    double hour_ms[48] = {0};  // [hour of day][0 = hour kernels, 1 = commit]
    uint64_t hour_launches[48] = {0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
    uint64_t launches = 0;
    // multi-region (Epidemiology::run_multi_engine): travel plan row of this region, host bookkeeping of the exchange
    uint32_t population = 0;  // live agents (P.n = slots)
    bool multi = false;
    int n_regions = 1;
    bool migration_enabled = false, commute_enabled = false;
    std::vector<uint32_t> migration_row, commute_row;  // [to]
    std::vector<uint32_t> migration_mat, commute_mat;  // the whole plan [from][to] (empty when disabled): sizes of the exchange segments
    std::shared_ptr<epi::Comm> comm;                   // epi_comm_init / epi_comm_init_local
    // TravelCounter (listeners/travel_counter.rs): outgoing migrators by destination and state, when asked for (epi_count_outgoing)
    bool count_outgoing = false;
    uint8_t* h_outgoing = nullptr;  // pinned copy of the send segments of a migration exchange
    size_t h_outgoing_bytes = 0;
    uint32_t outgoing_staged_hour = 0, outgoing_staged_stride = 0;
    std::vector<epi_outgoing_travel> outgoing_travels;
    uint32_t start_migration_hour = 0, end_migration_hour = 0;
    // The reference's sequential bookkeeping of the exchange lives on the device (travel.cu): the free-slot stack (LIFO: arrivals
    // pop, departures push) and the house / office occupancy heaps (grid.rs:47-80, 279-341) as occupancy arrays in tie order.
    // The host keeps the stack height and the initial images for epi_reset.
    bool pack_unsettled = false, unpack_unsettled = false;  // a deferred epi_travel_pack / unpack is in flight: the host's population mirror is stale
    unsigned travel_blocks = 0, travel_blocks_small = 0;  // grids of the cooperative exchange kernels (4 / 1 blocks per SM)
    std::vector<uint32_t> free_stack0, occ_house0, occ_office0;
    uint32_t* i_reg = nullptr;
    epi::TravelPtrs T{};
    uint32_t* t_block_counts = nullptr;
    std::vector<void*> travel_allocs;  // everything T points to (freed by epi_destroy)
    epi::TravelVars* h_tv = nullptr;   // pinned mirror of T.tv
    uint32_t* h_small = nullptr;       // pinned, 64 words
    epi_counts last_counts{};
    bool have_last_row = false;  // last_counts is the row of the hour just before the next one to run
    // host side of CitizenLocationMap::process_interventions (allocation_map.rs:306-337): the decisions
    epi::Interventions interventions;
    std::vector<epi_intervention_event> events;
    mutable std::string err;
};

namespace epi {
constexpr uint32_t RING_ROWS = 2400;  // counts ring: up to 100 simulated days between host synchronisations
int engine_fail(const epi_engine* e, int code, const std::string& msg);
void set_global_error(const std::string& msg);
int sync_travel(epi_engine* e);                                          // travel.cpp
int travel_status(epi_engine* e, uint32_t err, uint32_t population);  // travel.cpp
}  // namespace epi
int epi_travel_unpack_impl(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, uint64_t stride_records, uint32_t* counts_in, const uint32_t* wait_flags,
                           uint32_t exchange_no);
int epi_travel_pack_impl(epi_engine* e, uint32_t hour, int kind, void* send_buf, uint64_t stride_records, uint32_t* counts_out, epi::TravelRecord* const* peer_recv,
                         uint32_t* const* peer_flags, uint32_t exchange_no);
int epi_travel_exchange_fused(epi_engine* e, uint32_t hour, int kind, void* send_buf, const void* recv_buf, uint64_t stride_records, epi::TravelRecord* const* peer_recv,
                              uint32_t* const* peer_flags, const uint32_t* wait_flags, uint32_t exchange_no);
int epi_travel_unpack_wait(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, uint64_t stride_records, const uint32_t* wait_flags, uint32_t exchange_no);
namespace epi {
}  // namespace epi
