// One region engine: owns the HBM-resident state of a region and a CUDA stream; implements the C ABI of include/epi.h.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/epi.h"
#include "host_model.h"
#include "layout.h"
#include "simulation.h"

namespace epi {
// A set of small integers with find-first: 64-ary summary tree over a bitmap.
class RankSet {
  public:
    void init(size_t n) {
        levels_.clear();
        size_t words = (n + 63) / 64;
        for (;;) {
            levels_.emplace_back(words ? words : 1, 0ull);
            if (words <= 1) break;
            words = (words + 63) / 64;
        }
    }
    void set(uint32_t i) {
        for (auto& lv : levels_) {
            lv[i >> 6] |= 1ull << (i & 63);
            i >>= 6;
        }
    }
    void clear(uint32_t i) {
        for (auto& lv : levels_) {
            lv[i >> 6] &= ~(1ull << (i & 63));
            if (lv[i >> 6]) break;  // the summary bit above stays set
            i >>= 6;
        }
    }
    bool any() const { return levels_.back()[0] != 0; }
    uint32_t first() const {  // precondition: any()
        uint32_t i = 0;
        for (size_t l = levels_.size(); l-- > 0;) i = (i << 6) | (uint32_t)__builtin_ctzll(levels_[l][i]);
        return i;
    }

  private:
    std::vector<std::vector<uint64_t>> levels_;
};

// Grid::houses_occupancy / offices_occupancy (engine/src/geography/grid.rs:47-80, 279-341): a priority queue that pops the
// LEAST occupied area; ties go to the greatest Area in derive(Ord) order, i.e. the greatest (start.x, start.y)
// (grid.rs:67-73).  Occupancies are tiny (<= 4 per house, <= 100 per office), so this is a bucket queue: one RankSet per
// occupancy level over the areas' ranks in tie-break order -- O(1) per operation where a tree over 6 M houses costs
// microseconds of cache misses.
class OccupancyHeap {
  public:
    // start_xy[i] = (x, y) of area i's start_offset; max_level = capacity of an area
    void init(const std::vector<std::pair<int, int>>& start_xy, uint32_t max_level) {
        const size_t n = start_xy.size();
        occ_.assign(n, 0);
        present_.assign(n, 0);
        area_of_rank_.resize(n);
        for (size_t i = 0; i < n; ++i) area_of_rank_[i] = (uint32_t)i;
        std::sort(area_of_rank_.begin(), area_of_rank_.end(), [&](uint32_t a, uint32_t b) { return start_xy[a] > start_xy[b]; });  // greatest first
        rank_of_area_.resize(n);
        for (size_t r = 0; r < n; ++r) rank_of_area_[area_of_rank_[r]] = (uint32_t)r;
        levels_.assign(max_level + 2, RankSet());
        for (auto& l : levels_) l.init(n);
        lowest_ = 0;
    }
    void push(uint32_t i, uint32_t occupants) {
        occ_[i] = occupants;
        present_[i] = 1;
        level(occupants).set(rank_of_area_[i]);
        lowest_ = std::min(lowest_, occupants);
    }
    bool empty() {
        while (lowest_ < levels_.size() && !levels_[lowest_].any()) ++lowest_;
        return lowest_ >= levels_.size();
    }
    uint32_t pop_min() {  // BinaryHeap::pop; precondition: !empty()
        empty();
        const uint32_t i = area_of_rank_[levels_[lowest_].first()];
        levels_[lowest_].clear(rank_of_area_[i]);
        return i;
    }
    uint32_t occupants(uint32_t i) const { return occ_[i]; }
    void add_occupant(uint32_t i) {  // add_house_occupant / add_office_occupant after a pop
        occ_[i] += 1;
        level(occ_[i]).set(rank_of_area_[i]);
        lowest_ = std::min(lowest_, occ_[i]);
    }
    bool remove_occupant(uint32_t i) {  // remove_house_occupant / remove_office_occupant
        if (i >= present_.size() || !present_[i] || occ_[i] == 0) return false;
        level(occ_[i]).clear(rank_of_area_[i]);
        occ_[i] -= 1;
        level(occ_[i]).set(rank_of_area_[i]);
        lowest_ = std::min(lowest_, occ_[i]);
        return true;
    }

  private:
    RankSet& level(uint32_t occupants) { return levels_[std::min<size_t>(occupants, levels_.size() - 1)]; }
    std::vector<RankSet> levels_;
    std::vector<uint32_t> occ_, area_of_rank_, rank_of_area_;
    std::vector<uint8_t> present_;
    uint32_t lowest_ = 0;
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
