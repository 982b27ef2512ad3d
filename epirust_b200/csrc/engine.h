// One region engine: owns the HBM-resident state of a region and a CUDA stream; implements the C ABI of include/epi.h.
#pragma once
#include <cuda_runtime.h>

#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/epi.h"
#include "host_model.h"
#include "layout.h"
#include "simulation.h"

namespace epi {
// Grid::houses_occupancy / offices_occupancy (engine/src/geography/grid.rs:47-80, 279-341): pops the LEAST occupied area; ties
// go to the greatest Area in derive(Ord) order, i.e. the greatest (start.x, start.y) (grid.rs:67-73).
class OccupancyHeap {
  public:
    void init(size_t n_areas) { occ_.assign(n_areas, 0); present_.assign(n_areas, 0); x_.assign(n_areas, 0); y_.assign(n_areas, 0); q_.clear(); }
    void push(uint32_t i, uint32_t occupants, int start_x, int start_y) {
        occ_[i] = occupants; present_[i] = 1; x_[i] = start_x; y_[i] = start_y;
        q_.insert(key(i));
    }
    bool empty() const { return q_.empty(); }
    uint32_t pop_min() {  // BinaryHeap::pop
        auto it = q_.begin();
        const uint32_t i = std::get<3>(*it);
        q_.erase(it);
        return i;
    }
    uint32_t occupants(uint32_t i) const { return occ_[i]; }
    void add_occupant(uint32_t i) { occ_[i] += 1; q_.insert(key(i)); }  // add_house_occupant / add_office_occupant after a pop
    bool remove_occupant(uint32_t i) {                                   // remove_house_occupant / remove_office_occupant
        if (i >= present_.size() || !present_[i] || occ_[i] == 0) return false;
        q_.erase(key(i));
        occ_[i] -= 1;
        q_.insert(key(i));
        return true;
    }

  private:
    std::tuple<uint32_t, int, int, uint32_t> key(uint32_t i) const { return {occ_[i], -x_[i], -y_[i], i}; }
    std::set<std::tuple<uint32_t, int, int, uint32_t>> q_;
    std::vector<uint32_t> occ_;
    std::vector<uint8_t> present_;
    std::vector<int> x_, y_;
};
}  // namespace epi

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
    uint8_t* grid_alloc = nullptr;  // D.grid points pitch + GRID_XPAD bytes into this
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
    // day graph (hours h%24 = 1..23,0)
    cudaGraphExec_t day_graph = nullptr;
    uint32_t day_graph_launches = 0;
    bool graphs_enabled = true;
    // measurement
    bool timing = false;
    double kernel_ms[EPI_N_KERNEL_KINDS] = {0};
    uint64_t kernel_launches[EPI_N_KERNEL_KINDS] = {0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
    uint64_t launches = 0;
    // multi-region (Epidemiology::run_multi_engine): travel plan row of this region, host bookkeeping of the exchange
    uint32_t population = 0;  // live agents (P.n = slots)
    bool multi = false;
    int n_regions = 1;
    bool migration_enabled = false, commute_enabled = false;
    std::vector<uint32_t> migration_row, commute_row;  // [to]
    uint32_t start_migration_hour = 0, end_migration_hour = 0;
    std::vector<uint32_t> free_slots, free_slots0;     // LIFO: arrivals pop, departures push
    epi::OccupancyHeap houses_occupancy, offices_occupancy;
    std::vector<uint32_t> house_count0, office_count0;  // initial occupancies (epi_reset)
    uint32_t* i_reg = nullptr;
    // travel scratch on the device
    uint32_t *t_block_counts = nullptr, *t_total = nullptr, *t_out_slots = nullptr, *t_out_dest = nullptr, *t_idx = nullptr;
    uint32_t *t_table_keys = nullptr, *t_table_vals = nullptr;
    uint8_t* t_placed = nullptr;
    size_t t_list_capacity = 0, t_table_capacity = 0;
    uint32_t* h_small = nullptr;  // pinned, 64 words
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
}  // namespace epi
