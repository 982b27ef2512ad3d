// =============================================================================
// oracle/epi_oracle_travel.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Restates the reference's multi-engine hour loop and traveller exchange
// (engine/src/epidemiology_simulation.rs:276-547, engine/src/travel/**,
// engine/src/allocation_map.rs:165-303, engine/src/geography/grid.rs:279-341,
// engine/src/citizen/citizen_factory.rs:90-110) for R regions held in ONE
// process: the MPI / Kafka transport (engine/src/transport/*.rs) carries no
// arithmetic, so "send" and "receive" are vector moves here.
//
// PARITY STATUS: "parity unpinned" like the rest of the stochastic path (see
// epi_oracle.hpp).  The reference leaves four orders to chance (hash-map
// iteration, MPI wait_any arrival order, HashMap-built heaps, reservoir
// sampling); the deterministic conventions chosen instead -- and shared with
// the CUDA path so that multi-region runs are bit-identical oracle <-> GPU --
// are stated at each site below and in DESIGN.md "Traveller exchange".
// =============================================================================
#pragma once
#include <set>
#include <tuple>

#include "epi_oracle_engine.hpp"

namespace orc {

// common/src/config/travel_plan_config.rs:22-41 with regions = indices 0..R-1
struct TravelPlanConfig {
    int n_regions = 0;
    bool migration_enabled = false, commute_enabled = false;
    std::vector<uint32_t> migration, commute;  // R x R row-major: [from][to]
    Hour start_migration_hour = 0, end_migration_hour = 0;
    // common/src/models/travel_plan.rs:28-49
    uint32_t get_outgoing(const std::vector<uint32_t>& m, int from, int to) const { return m[(size_t)from * n_regions + to]; }
    uint32_t get_total_outgoing(const std::vector<uint32_t>& m, int from) const {
        uint32_t s = 0;
        for (int to = 0; to < n_regions; ++to) s += get_outgoing(m, from, to);
        return s;
    }
};

// travel/migration/migrator.rs:25-32, travel/commute/commuter.rs:26-35 (id = the sender's slot, informational)
struct Migrator {
    uint32_t id; int immunity; bool vaccinated, uses_public_transport, working; DiseaseStateMachine state_machine;
};
struct Commuter {
    uint32_t id; int immunity; Area home_location, work_location; bool vaccinated, uses_public_transport, working; DiseaseStateMachine state_machine;
};

// Grid::houses_occupancy / offices_occupancy (grid.rs:47-80, 279-341): a priority queue that pops the LEAST occupied
// area; ties go to the GREATEST Area in derive(Ord) order = (location_id, start.x, start.y, ...) (grid.rs:67-73).
struct OccupancyHeap {
    // key: (occupants, -start.x, -start.y) ascending == the reference's pop order
    std::set<std::tuple<uint32_t, int, int, uint32_t>> q;  // + index of the area in its list
    std::vector<uint32_t> occ;
    std::vector<uint8_t> present;
    const std::vector<Area>* areas = nullptr;
    void init(const std::vector<Area>& a) { areas = &a; occ.assign(a.size(), 0); present.assign(a.size(), 0); q.clear(); }
    std::tuple<uint32_t, int, int, uint32_t> key(uint32_t i) const { return {occ[i], -(*areas)[i].start_offset.x, -(*areas)[i].start_offset.y, i}; }
    void push(uint32_t i, uint32_t occupants) { occ[i] = occupants; present[i] = 1; q.insert(key(i)); }
    uint32_t pop_min() {  // BinaryHeap::pop
        if (q.empty()) throw std::runtime_error("occupancy heap is empty");
        auto it = q.begin();
        uint32_t i = std::get<3>(*it);
        q.erase(it);
        return i;
    }
    void add_occupant(uint32_t i) { occ[i] += 1; q.insert(key(i)); }  // add_house_occupant / add_office_occupant
    void remove_occupant(uint32_t i) {                                 // remove_*_occupant: panics when the area is not in the heap
        if (!present[i]) throw std::runtime_error("Could not find house / office");
        q.erase(key(i));
        occ[i] -= 1;
        q.insert(key(i));
    }
};

struct RegionEngine : Engine {
    const TravelPlanConfig* plan = nullptr;
    uint32_t capacity = 0;
    std::vector<uint32_t> free_slots;  // LIFO; arrivals pop, departures push (convention: see MultiEngine::exchange)
    OccupancyHeap houses_occupancy, offices_occupancy;
    std::vector<std::pair<Point, Migrator>> outgoing_migrators;
    std::vector<std::pair<Point, Commuter>> outgoing_commuters;
    Hour exchanges = 0;

    // Citizen::can_migrate (citizen/mod.rs:456-466)
    bool can_migrate(const Citizen& z, Hour hour) const {
        return hour % 24 == 0 && hour > plan->start_migration_hour && hour < plan->end_migration_hour && z.work_location.location_id == region &&
               z.home_location.location_id == region && z.can_move();
    }
    // Citizen::is_commuter (citizen/mod.rs:488-495)
    bool is_commuter(const Citizen& z, Hour hour) const {
        return (hour % 24 == constants::ROUTINE_TRAVEL_START_TIME && z.can_move() && z.work_location.location_id != region) ||
               (hour % 24 == constants::ROUTINE_TRAVEL_END_TIME && z.can_move() && z.home_location.location_id != region);
    }

    void init_region(const orc_config& c, uint64_t seed_, int region_, const TravelPlanConfig* plan_, uint32_t extra_capacity, int threads_) {
        region = region_;
        plan = plan_;
        Engine::init(c, seed_, Rng::KEYED, threads_);
        capacity = cfg.number_of_agents + extra_capacity;
        free_slots.clear();
        for (uint32_t s = capacity; s-- > cfg.number_of_agents;) free_slots.push_back(s);  // pop order: n, n+1, ...
        // update_commuters (citizen_factory.rs:90-110): the first sum(row) working public-transport users in creation order
        // get the regions of the commute row as work region, row order, `count` agents each
        std::vector<Citizen*> by_id(cfg.number_of_agents, nullptr);
        PointMap& m = map.current_locations;
        for (size_t i = 0; i < m.capacity(); ++i) if (m.used[i]) by_id[m.vals[i].id] = &m.vals[i];
        if (plan && plan->commute_enabled) {
            uint32_t a = 0;
            for (int to = 0; to < plan->n_regions; ++to) {
                uint32_t want = plan->get_outgoing(plan->commute, region, to);
                while (want > 0 && a < cfg.number_of_agents) {
                    Citizen& z = *by_id[a++];
                    if (z.is_working() && z.work_location.location_id == region && z.uses_public_transport) {
                        z.work_location.location_id = to;
                        --want;
                    }
                }
            }
        }
        // set_start_locations_and_occupancies (grid.rs:125-155): houses that have residents enter the heap with their
        // resident count; every office enters with its number of workers whose work region is this one (:262-277)
        houses_occupancy.init(map.grid.houses);
        offices_occupancy.init(map.grid.offices);
        std::vector<uint32_t> hc(map.grid.houses.size(), 0), oc(map.grid.offices.size(), 0);
        for (uint32_t a = 0; a < cfg.number_of_agents; ++a) {
            const Citizen& z = *by_id[a];
            hc[house_index(map.grid, z.home_location)]++;
            if (z.is_working() && z.work_location.location_id == region) oc[office_index(map.grid, z.work_location)]++;
        }
        for (uint32_t i = 0; i < hc.size(); ++i) if (hc[i] > 0) houses_occupancy.push(i, hc[i]);
        for (uint32_t i = 0; i < oc.size(); ++i) offices_occupancy.push(i, oc[i]);
    }

    // CitizenLocationMap::simulate with travel_plan_config = Some(..) (allocation_map.rs:67-129): the standalone step
    // plus traveller selection in phase-B order (here: ascending slot id).
    void simulate_travel(Hour hour, double percent_outgoing) {
        outgoing_migrators.clear();
        outgoing_commuters.clear();
        Engine::simulate(counts_at_hr, hour, nullptr);
        // `updates` still holds this hour's phase-A results in ascending id; replay the selection over it
        for (const Update& u : updates) {
            const Citizen* now = nullptr;  // where did the agent end up (allocation_map.rs:96-101)
            Point new_location = u.new_cell;
            now = map.current_locations.get(u.new_cell);
            if (!now || now->id != u.agent.id) new_location = u.old_cell;
            const Citizen& agent = u.agent;
            if (plan->migration_enabled && can_migrate(agent, hour)) {
                Rng r = engine_rng(agent.id, hour, DOM_MIGRATE);
                if (r.gen_bool(0, percent_outgoing))
                    outgoing_migrators.push_back({new_location, Migrator{agent.id, agent.immunity, agent.vaccinated, agent.uses_public_transport, agent.is_working(), agent.state_machine}});
            }
            if (plan->commute_enabled && is_commuter(agent, hour))
                outgoing_commuters.push_back({new_location, Commuter{agent.id, agent.immunity, agent.home_location, agent.work_location, agent.vaccinated,
                                                                      agent.uses_public_transport, agent.is_working(), agent.state_machine}});
        }
    }

    static void decrement_counts(const State& s, Counts& c) {  // allocation_map.rs:291-301
        switch (s.kind) {
            case Susceptible: c.susceptible -= 1; break;
            case Exposed: c.exposed -= 1; break;
            case Infected: c.infected -= 1; break;
            case Recovered: c.recovered -= 1; break;
            default: throw std::runtime_error("Deceased agent should not travel!");
        }
    }
    static void increment_counts(const State& s, Counts& c) {  // allocation_map.rs:279-289
        switch (s.kind) {
            case Susceptible: c.susceptible += 1; break;
            case Exposed: c.exposed += 1; break;
            case Infected: c.infected += 1; break;
            case Recovered: c.recovered += 1; break;
            default: throw std::runtime_error("Should not receive deceased agent!");
        }
    }

    // remove_migrators (allocation_map.rs:165-192)
    void remove_migrators(const std::vector<std::pair<Point, Migrator>>& outgoing) {
        for (auto& pm : outgoing) {
            decrement_counts(pm.second.state_machine.state, counts_at_hr);
            Citizen z;
            if (!map.current_locations.remove(pm.first, &z)) throw std::runtime_error("Trying to remove citizen from a location where no citizen is present");
            houses_occupancy.remove_occupant(house_index(map.grid, z.home_location));
            if (z.is_working()) offices_occupancy.remove_occupant(office_index(map.grid, z.work_location));
            free_slots.push_back(z.id);
        }
    }
    // remove_commuters (allocation_map.rs:194-212)
    void remove_commuters(const std::vector<std::pair<Point, Commuter>>& outgoing) {
        for (auto& pc : outgoing) {
            decrement_counts(pc.second.state_machine.state, counts_at_hr);
            Citizen z;
            if (!map.current_locations.remove(pc.first, &z)) throw std::runtime_error("Trying to remove citizen from a location where no citizen is present");
            free_slots.push_back(z.id);
        }
    }
    uint32_t take_slot() {
        if (free_slots.empty()) throw std::runtime_error("region is out of agent slots (raise extra_capacity)");
        uint32_t s = free_slots.back();
        free_slots.pop_back();
        return s;
    }

    // select_starting_points (allocation_map.rs:339-347): `n` distinct vacant cells of area, x in [sx, ex), y in [sy, ey).
    // The reference reservoir-samples the vacant cells; convention here (and in the CUDA path): rounds.  In round a every
    // still-unplaced arrival k walks its own candidate sequence Philox(seed, k, hour, DOM_ARRIVAL) blocks
    // a * PLACE_TRIES .. a * PLACE_TRIES + PLACE_TRIES - 1 and proposes the FIRST candidate that is vacant now (start-of-hour
    // occupants and earlier rounds' winners count as occupied); of several proposals for one cell the lowest k wins; a
    // loser, or an arrival whose candidates were all occupied, tries again in the next round.
    static constexpr uint32_t PLACE_TRIES = 8;
    uint32_t max_place_rounds = 0;  // most rounds one select_starting_points call needed so far (test hook: orc_multi_max_place_rounds)
    std::vector<Point> select_starting_points(const Area& area, size_t n, Hour hour) {
        std::vector<Point> out(n);
        std::vector<uint8_t> placed(n, 0);
        PointMap taken; taken.init(n);
        const uint32_t w = (uint32_t)(area.end_offset.x - area.start_offset.x), h = (uint32_t)(area.end_offset.y - area.start_offset.y);
        size_t left = n;
        for (uint32_t attempt = 0; left > 0; ++attempt) {
            if (attempt > 64) throw std::runtime_error("Not enough locations are available for travellers");
            PointMap round; round.init(left);
            std::vector<Point> prop(n);
            std::vector<uint8_t> has_prop(n, 0);
            for (size_t k = 0; k < n; ++k) {
                if (placed[k]) continue;
                Rng r; r.mode = Rng::KEYED; r.seed = seed; r.agent = (uint32_t)k; r.hour = hour; r.domain = DOM_ARRIVAL;
                for (uint32_t t = 0; t < PLACE_TRIES; ++t) {
                    uint32_t o[4];
                    r.block(attempt * PLACE_TRIES + t, o);
                    Point p{area.start_offset.x + (int)mulhi32(o[0], w), area.start_offset.y + (int)mulhi32(o[1], h)};
                    if (!map.is_cell_vacant(p) || taken.contains_key(p)) continue;
                    prop[k] = p;
                    has_prop[k] = 1;
                    break;
                }
                if (!has_prop[k]) continue;
                Citizen marker; marker.id = (uint32_t)k;
                Citizen& first = round.entry_or_insert(prop[k], marker);  // ascending k: the first entry is the lowest k
                (void)first;
            }
            for (size_t k = 0; k < n; ++k) {
                if (placed[k] || !has_prop[k]) continue;
                const Citizen* win = round.get(prop[k]);
                if (win && win->id == (uint32_t)k) { out[k] = prop[k]; placed[k] = 1; --left; Citizen m; taken.insert(prop[k], m); }
            }
            max_place_rounds = std::max(max_place_rounds, attempt + 1u);
        }
        return out;
    }

    // assimilate_migrators (allocation_map.rs:214-243)
    void assimilate_migrators(const std::vector<Migrator>& incoming, Hour hour) {
        if (incoming.empty()) return;
        std::vector<Point> locations = select_starting_points(map.grid.housing_area, incoming.size(), hour);
        for (size_t k = 0; k < incoming.size(); ++k) {
            const Migrator& mg = incoming[k];
            const uint32_t house = houses_occupancy.pop_min();
            if (houses_occupancy.occ[house] >= constants::HOME_SIZE * constants::HOME_SIZE) throw std::runtime_error("Couldn't find any house with free space!");
            uint32_t office = 0;
            if (mg.working) {
                office = offices_occupancy.pop_min();
                if (offices_occupancy.occ[office] >= constants::OFFICE_SIZE * constants::OFFICE_SIZE) throw std::runtime_error("Couldn't find any offices with free space!");
            }
            Citizen z;  // Citizen::from_migrator (citizen/mod.rs:113-136)
            z.id = take_slot();
            z.immunity = mg.immunity;
            z.home_location = map.grid.houses[house];
            z.work_location = mg.working ? map.grid.offices[office] : map.grid.houses[house];
            z.vaccinated = mg.vaccinated;
            z.uses_public_transport = mg.uses_public_transport;
            z.hospitalized = false;
            z.transport_location = locations[k];
            z.state_machine = mg.state_machine;
            z.isolated = false;
            z.current_area = map.grid.housing_area;
            z.work_status = NA;
            z.work_quarantined = false;
            houses_occupancy.add_occupant(house);
            if (mg.working) offices_occupancy.add_occupant(office);
            increment_counts(z.state_machine.state, counts_at_hr);
            if (map.current_locations.insert(locations[k], z)) throw std::runtime_error("assert!(result.is_none()) failed");
        }
    }
    // assimilate_commuters (allocation_map.rs:245-277)
    void assimilate_commuters(const std::vector<Commuter>& incoming, Hour hour) {
        if (incoming.empty()) return;
        std::vector<Point> locations = select_starting_points(map.grid.transport_area, incoming.size(), hour);
        for (size_t k = 0; k < incoming.size(); ++k) {
            const Commuter& cm = incoming[k];
            Citizen z;  // Citizen::from_commuter (citizen/mod.rs:138-154)
            z.id = take_slot();
            z.immunity = cm.immunity;
            z.home_location = cm.home_location;
            z.work_location = cm.work_location;
            if (hour == constants::ROUTINE_TRAVEL_START_TIME) {  // sic: the absolute hour 7, i.e. the first day only (:260)
                const uint32_t office = offices_occupancy.pop_min();
                if (offices_occupancy.occ[office] >= constants::OFFICE_SIZE * constants::OFFICE_SIZE) throw std::runtime_error("Couldn't find any offices with free space!");
                offices_occupancy.add_occupant(office);
                z.work_location = map.grid.offices[office];
            }
            z.vaccinated = cm.vaccinated;
            z.uses_public_transport = cm.uses_public_transport;
            z.hospitalized = false;
            z.transport_location = locations[k];
            z.state_machine = cm.state_machine;
            z.isolated = false;
            z.current_area = map.grid.housing_area;
            z.work_status = Normal;
            z.work_quarantined = false;
            increment_counts(z.state_machine.state, counts_at_hr);
            if (map.current_locations.insert(locations[k], z)) throw std::runtime_error("assert!(result.is_none()) failed");
        }
    }
};

// R engines in lock step: Epidemiology::run_multi_engine for every region, with the transport replaced by vector moves.
struct MultiEngine {
    TravelPlanConfig plan;
    std::vector<RegionEngine> regions;

    void init(const std::vector<orc_config>& cfgs, uint64_t seed, const TravelPlanConfig& p, uint32_t extra_capacity, int threads) {
        plan = p;
        regions.resize(cfgs.size());
        for (size_t r = 0; r < cfgs.size(); ++r) regions[r].init_region(cfgs[r], seed + r, (int)r, &plan, extra_capacity, threads);
    }

    // one simulated hour of every region (epidemiology_simulation.rs:331-537)
    void step(Hour hour) {
        const int R = (int)regions.size();
        const Hour h = hour % 24;
        const bool migration_hour = plan.migration_enabled && h == 0;
        const bool commute_hour = plan.commute_enabled && (h == constants::ROUTINE_TRAVEL_START_TIME || h == constants::ROUTINE_TRAVEL_END_TIME);
        // outgoing per (from, to)
        std::vector<std::vector<Migrator>> mig((size_t)R * R);
        std::vector<std::vector<Commuter>> com((size_t)R * R);
        std::vector<std::vector<std::pair<Point, Migrator>>> actual_outgoing((size_t)R);
        for (int r = 0; r < R; ++r) {
            RegionEngine& e = regions[(size_t)r];
            e.counts_at_hr.hour = hour;
            const Count pop = e.map.current_population();
            if (pop == 0) throw std::runtime_error("No citizens!");
            double percent_outgoing = 0.0;  // EngineMigrationPlan::percent_outgoing (engine_migration_plan.rs:44-49)
            if (migration_hour) percent_outgoing = (double)plan.get_total_outgoing(plan.migration, r) / (double)pop;
            e.simulate_travel(hour, percent_outgoing);
            if (plan.migration_enabled) {
                // alloc_outgoing_to_regions (engine_migration_plan.rs:51-77) + MigratorsByRegion::alloc_citizens
                // (migrators_by_engine.rs:34-55): regions in plan order take floor(share * total) from the front
                const size_t total = e.outgoing_migrators.size();
                const uint32_t planned_total = plan.get_total_outgoing(plan.migration, r);
                size_t front = 0;
                for (int to = 0; to < R; ++to) {
                    if (to == r || plan.get_outgoing(plan.migration, r, to) == 0) continue;
                    const double share = (double)plan.get_outgoing(plan.migration, r, to) / (double)planned_total;
                    size_t count = (size_t)(int32_t)(share * (double)(int32_t)total);
                    if (count > total - front) count = total - front;
                    for (size_t k = 0; k < count; ++k) {
                        mig[(size_t)r * R + to].push_back(e.outgoing_migrators[front + k].second);
                        actual_outgoing[(size_t)r].push_back(e.outgoing_migrators[front + k]);
                    }
                    front += count;
                }
            }
            if (commute_hour) {  // CommutersByRegion::get_commuters_by_region (commuters_by_region.rs:59-78)
                for (auto& pc : e.outgoing_commuters) {
                    const int to = h == constants::ROUTINE_TRAVEL_START_TIME ? pc.second.work_location.location_id : pc.second.home_location.location_id;
                    com[(size_t)r * R + to].push_back(pc.second);
                }
            }
        }
        for (int r = 0; r < R; ++r) {
            RegionEngine& e = regions[(size_t)r];
            if (plan.migration_enabled) {
                // receive order convention: source region index ascending (the reference takes MPI wait_any order)
                std::vector<Migrator> incoming;
                if (migration_hour)
                    for (int from = 0; from < R; ++from) incoming.insert(incoming.end(), mig[(size_t)from * R + r].begin(), mig[(size_t)from * R + r].end());
                e.remove_migrators(migration_hour ? actual_outgoing[(size_t)r] : std::vector<std::pair<Point, Migrator>>());
                e.assimilate_migrators(incoming, hour);
            }
            if (commute_hour) {
                std::vector<Commuter> incoming;
                for (int from = 0; from < R; ++from) incoming.insert(incoming.end(), com[(size_t)from * R + r].begin(), com[(size_t)from * R + r].end());
                e.remove_commuters(e.outgoing_commuters);
                e.assimilate_commuters(incoming, hour);
            }
            // listeners.counts_updated; process_interventions; stop_simulation's MultiEngine arm (:564-571)
            e.process_interventions();
            const Counts& c = e.counts_at_hr;
            if (e.interventions.lockdown.is_locked_down && c.exposed == 0 && c.infected == 0 && c.hospitalized == 0)
                e.interventions.lockdown.set_zero_infection_hour(c.hour);
        }
    }
};

}  // namespace orc
