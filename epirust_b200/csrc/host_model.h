// Host-side model set-up of one region: geography and the Auto population factory, producing the structure-of-arrays
// image that is uploaded to HBM once.  Product code (the oracle has its own independent restatement).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/epi.h"
#include "layout.h"

namespace epi {

struct Geometry {
    int grid_size = 0;
    Rect housing{}, transport{}, work{}, hospital_initial{}, hospital_resized{}, hospital_expanded{};
    int house_nx = 0, house_ny = 0, office_nx = 0, office_ny = 0;
    uint32_t n_houses = 0, n_offices = 0;
    uint32_t pitch = 0, rows = 0;
};

// geography::define_geography + Grid::resize_hospital (geography/mod.rs:33-70, grid.rs:240-261)
Geometry make_geometry(uint32_t grid_size, uint32_t n_agents, double hospital_beds_percentage);

// house / office index (row-major in area_factory order, geography/area.rs:95-117) <-> packed origin cell
inline uint32_t house_origin(const Geometry& g, uint32_t idx) {
    return ((uint32_t)(g.housing.sy + 2 * (int)(idx / (uint32_t)g.house_nx)) << CELL_BITS) | (uint32_t)(g.housing.sx + 2 * (int)(idx % (uint32_t)g.house_nx));
}
inline uint32_t office_origin(const Geometry& g, uint32_t idx) {
    return ((uint32_t)(g.work.sy + 10 * (int)(idx / (uint32_t)g.office_nx)) << CELL_BITS) | (uint32_t)(g.work.sx + 10 * (int)(idx % (uint32_t)g.office_nx));
}
inline uint32_t house_index_of(const Geometry& g, uint32_t origin) {
    const int x = (int)(origin & CELL_XMASK), y = (int)((origin >> CELL_BITS) & CELL_XMASK);
    return (uint32_t)((y - g.housing.sy) / 2 * g.house_nx + (x - g.housing.sx) / 2);
}
inline uint32_t office_index_of(const Geometry& g, uint32_t origin) {
    const int x = (int)(origin & CELL_XMASK), y = (int)((origin >> CELL_BITS) & CELL_XMASK);
    return (uint32_t)((y - g.work.sy) / 10 * g.office_nx + (x - g.work.sx) / 10);
}

// how agent ids are assigned to the citizens of the population (see the "Agent numbering" note in host_model.cpp)
enum class AgentOrder { Creation, House };
AgentOrder agent_order();

struct HostAgents {
    std::vector<uint32_t> cell, st, t0, home, work, wsa, reg;
    size_t size() const { return st.size(); }
    void resize(size_t n) { cell.resize(n); st.resize(n); t0.resize(n); home.resize(n); work.resize(n); wsa.resize(n); reg.resize(n); }
};

// Population::Csv: the columns of the PopulationRecords the engine reads (citizen/population_record.rs:23-31, citizen/mod.rs:155-180),
// in file order == the reference's creation order
struct PopulationRecords {
    std::vector<uint8_t> working, pub_transport;
    size_t size() const { return working.size(); }
};
// csv::Reader::from_reader(file).deserialize::<PopulationRecord>() (grid.rs:202-208).  Throws std::runtime_error.
PopulationRecords read_population_csv(const std::string& path);
// `cfg` with number_of_agents resolved: the record count when cfg.population_csv_file is set (records are then loaded into `records`)
epi_config resolve_population(const epi_config& cfg, PopulationRecords& records);

// Grid::generate_population + citizen_factory + set_starting_infections + init_interventions' essential workers
// (grid.rs:83-155, citizen_factory.rs:31-134, epidemiology_simulation.rs:178-192).  Throws std::runtime_error.
// With `records` (Population::Csv, Grid::read_population grid.rs:194-231) working / uses_public_transport come from the file.
void build_population(const epi_config& cfg, const Geometry& geo, uint64_t seed, int region, HostAgents& out, const PopulationRecords* records = nullptr);

// citizen_factory::update_commuters (citizen_factory.rs:90-110): the first sum(commute_row) working public-transport users
// in creation order get the row's regions as work region (row order, commute_row[to] agents each)
void apply_commute_plan(HostAgents& agents, uint32_t n_agents, int region, const std::vector<uint32_t>& commute_row);

// Params for the kernels from config + geometry
Params make_params(const epi_config& cfg, const Geometry& geo, uint64_t seed, int region);

// validation of what the reference would panic on later (and our own packing limits)
std::string validate_config(const epi_config& cfg);

}  // namespace epi
