// =============================================================================
// oracle/epi_oracle_engine.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Restates the reference's population factory, CitizenLocationMap::simulate,
// interventions and the single-engine hour loop.  See epi_oracle.hpp for the
// parity status ("parity unpinned" for stochastic outcomes).
// =============================================================================
#pragma once
#include <omp.h>

#include <cstdlib>

#include <map>
#include <tuple>
#include <unordered_set>

#include "epi_oracle.hpp"

namespace orc {

// Plain-C mirror of common::config::Config (common/src/config/mod.rs:44-58) for the Auto population path.
// Field order is shared with include/epi.h `epi_config` so one ctypes struct serves both.
extern "C" struct orc_config {
    // population.Auto (population.rs:36-43)
    uint32_t number_of_agents;
    double public_transport_percentage;
    double working_percentage;
    // disease (disease/mod.rs:26-45)
    uint32_t regular_transmission_start_day, high_transmission_start_day, last_day;
    uint32_t asymptomatic_last_day, mild_infected_last_day;
    double regular_transmission_rate, high_transmission_rate, death_rate;
    double percentage_asymptomatic_population, percentage_severe_infected_population;
    uint32_t exposed_duration, pre_symptomatic_duration;
    // geography_parameters (geography_parameters.rs:24-28)
    uint32_t grid_size;
    double hospital_beds_percentage;
    uint32_t hours;
    // starting_infections (starting_infections.rs:22-28)
    uint32_t infected_mild_asymptomatic, infected_mild_symptomatic, infected_severe, exposed;
    // interventions (intervention_config.rs:23-48)
    int32_t has_lockdown;
    uint32_t lockdown_at_number_of_infections;
    double essential_workers_population;
    int32_t has_build_new_hospital;
    uint32_t spread_rate_threshold;
    int32_t n_vaccinations;
    uint32_t vaccinate_at_hour[8];
    double vaccinate_percent[8];
    // population.Csv.file (population.rs:30-34), "" for population.Auto
    char population_csv_file[256];
};

// PopulationRecord (citizen/population_record.rs:23-43) as read by csv::Reader::deserialize in Grid::read_population
// (grid.rs:202-208): header row, columns matched by name (ind: u32, age: String, working / pub_transport: "True" | "False").
struct PopulationRecord {
    uint32_t ind;
    std::string age;
    bool working, pub_transport;
};
inline std::vector<PopulationRecord> read_population_records(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Could not read population file");
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
    std::fclose(f);
    // tokenise: records end at \n (or \r\n), fields at ',', a field may be wrapped in double quotes ("" = a quote)
    std::vector<std::vector<std::string>> table;
    std::vector<std::string> row;
    std::string field;
    bool in_quotes = false, any = false;
    auto end_field = [&] { row.push_back(field); field.clear(); };
    auto end_row = [&] {
        end_field();
        if (!(row.size() == 1 && row[0].empty())) table.push_back(row);  // blank lines are skipped
        row.clear();
        any = false;
    };
    for (size_t i = 0; i < text.size(); ++i) {
        const char ch = text[i];
        if (in_quotes) {
            if (ch == '"' && i + 1 < text.size() && text[i + 1] == '"') { field.push_back('"'); ++i; }
            else if (ch == '"') in_quotes = false;
            else field.push_back(ch);
            continue;
        }
        if (ch == '"' && field.empty()) { in_quotes = true; any = true; }
        else if (ch == ',') { end_field(); any = true; }
        else if (ch == '\n') end_row();
        else if (ch == '\r' && i + 1 < text.size() && text[i + 1] == '\n') continue;
        else { field.push_back(ch); any = true; }
    }
    if (in_quotes) throw std::runtime_error("Could not deserialize population line: unterminated quote");
    if (any || !field.empty() || !row.empty()) end_row();
    if (table.empty()) return {};
    const std::vector<std::string>& header = table[0];
    auto column = [&](const char* name) {
        for (size_t k = 0; k < header.size(); ++k) if (header[k] == name) return k;
        throw std::runtime_error(std::string("Could not deserialize population line: missing field `") + name + "`");
    };
    const size_t c_ind = column("ind"), c_age = column("age"), c_working = column("working"), c_pt = column("pub_transport");
    auto to_bool = [](const std::string& v) {
        if (v == "True") return true;
        if (v == "False") return false;
        throw std::runtime_error("Could not deserialize population line: expected True or False, got \"" + v + "\"");
    };
    std::vector<PopulationRecord> out;
    for (size_t r = 1; r < table.size(); ++r) {
        const std::vector<std::string>& rec = table[r];
        if (rec.size() != header.size()) throw std::runtime_error("Could not deserialize population line: unequal record lengths");
        PopulationRecord pr;
        const std::string& ind = rec[c_ind];
        if (ind.empty() || ind.size() > 10) throw std::runtime_error("Could not deserialize population line: bad `ind`");
        uint64_t v = 0;
        for (char d : ind) { if (d < '0' || d > '9') throw std::runtime_error("Could not deserialize population line: bad `ind`"); v = v * 10 + (uint64_t)(d - '0'); }
        if (v > 0xFFFFFFFFull) throw std::runtime_error("Could not deserialize population line: bad `ind`");
        pr.ind = (uint32_t)v;
        pr.age = rec[c_age];
        pr.working = to_bool(rec[c_working]);
        pr.pub_transport = to_bool(rec[c_pt]);
        out.push_back(pr);
    }
    return out;
}

inline Disease disease_from(const orc_config& c) {
    Disease d;
    d.regular_transmission_start_day = c.regular_transmission_start_day;
    d.high_transmission_start_day = c.high_transmission_start_day;
    d.last_day = c.last_day;
    d.asymptomatic_last_day = c.asymptomatic_last_day;
    d.mild_infected_last_day = c.mild_infected_last_day;
    d.regular_transmission_rate = c.regular_transmission_rate;
    d.high_transmission_rate = c.high_transmission_rate;
    d.death_rate = c.death_rate;
    d.percentage_asymptomatic_population = c.percentage_asymptomatic_population;
    d.percentage_severe_infected_population = c.percentage_severe_infected_population;
    d.exposed_duration = c.exposed_duration;
    d.pre_symptomatic_duration = c.pre_symptomatic_duration;
    return d;
}

// -----------------------------------------------------------------------------
// interventions/lockdown.rs
// -----------------------------------------------------------------------------
struct LockdownIntervention {
    bool is_locked_down = false;
    bool has_config = false;
    Count at_number_of_infections = 0;
    double essential_workers_population = 0.0;
    Hour zero_infection_hour = 0;
    // :55-57
    bool should_apply(const Counts& c) const { return !is_locked_down && c.hour % constants::HOURS_IN_A_DAY == 0 && above_threshold(c); }
    // :59-61
    bool above_threshold(const Counts& c) const { return has_config && c.infected > at_number_of_infections; }
    // :63-67
    void set_zero_infection_hour(Hour h) { if (zero_infection_hour == 0) zero_infection_hour = h; }
    // :69-73
    bool should_unlock(const Counts& c) const {
        Hour unlock_hour = zero_infection_hour + (Hour)std::round((double)constants::QUARANTINE_DAYS * 1.5) * constants::HOURS_IN_A_DAY;
        return is_locked_down && c.hour == unlock_hour;
    }
    void apply() { if (!has_config) throw std::runtime_error("Tried to apply lockdown when intervention is not present"); is_locked_down = true; }  // :75-84
    void unapply() { is_locked_down = false; zero_infection_hour = 0; }  // :86-89
    double get_essential_workers_percentage() const { return has_config ? essential_workers_population : 0.0; }  // :91-96
};
// interventions/hospital.rs
struct BuildNewHospital {
    Count new_infections_in_a_day = 0;
    bool has_config = false;
    uint32_t spread_rate_threshold = 0;
    bool has_applied = false;
    // :51-58
    bool should_apply(const Counts& c) const {
        if (has_applied) return false;
        bool start_of_day = c.hour % 24 == 0;
        bool exceeds = has_config && new_infections_in_a_day >= spread_rate_threshold;
        return start_of_day && exceeds;
    }
    void apply() { has_applied = true; }
    // :68-74 (sic: saturating_sub of the previous value)
    void counts_updated(const Counts& c) {
        if (c.hour % 24 == 0) new_infections_in_a_day = c.infected >= new_infections_in_a_day ? c.infected - new_infections_in_a_day : 0;
    }
};
// interventions/vaccination.rs
struct VaccinateIntervention {
    std::map<Hour, Percentage> intervention;  // :27-29 HashMap<Hour, Percentage>; later entries overwrite (:43-45)
    bool get_vaccination_percentage(const Counts& c, double& p) const {  // :52-54
        auto it = intervention.find(c.hour);
        if (it == intervention.end()) return false;
        p = it->second;
        return true;
    }
};
struct Interventions {
    VaccinateIntervention vaccinate;
    LockdownIntervention lockdown;
    BuildNewHospital build_new_hospital;
};
inline Interventions interventions_from(const orc_config& c) {
    Interventions iv;
    for (int i = 0; i < c.n_vaccinations && i < 8; ++i) iv.vaccinate.intervention[c.vaccinate_at_hour[i]] = c.vaccinate_percent[i];
    iv.lockdown.has_config = c.has_lockdown != 0;
    iv.lockdown.at_number_of_infections = c.lockdown_at_number_of_infections;
    iv.lockdown.essential_workers_population = c.essential_workers_population;
    iv.build_new_hospital.has_config = c.has_build_new_hospital != 0;
    iv.build_new_hospital.spread_rate_threshold = c.spread_rate_threshold;
    return iv;
}

struct InterventionEvent {  // listeners/intervention_reporter.rs:28-33
    Hour hour;
    int kind;    // 0 lockdown, 1 vaccination, 2 build_new_hospital
    int status;  // lockdown: 1 locked_down, 0 lockdown_revoked ; else 0
};

// -----------------------------------------------------------------------------
// The engine: Epidemiology (epidemiology_simulation.rs) for RunMode::Standalone
// -----------------------------------------------------------------------------
struct Engine {
    orc_config cfg;
    std::vector<PopulationRecord> records;  // Population::Csv only
    std::vector<uint32_t> creation_of;      // agent id -> the reference's creation number (init draws are keyed on it)
    Disease disease;
    CitizenLocationMap map;
    Counts counts_at_hr;
    Interventions interventions;
    std::vector<InterventionEvent> events;
    uint64_t seed = 0;
    Rng::Mode mode = Rng::KEYED;
    int threads = 1;
    bool shuffle_phase_b = false;
    std::vector<std::mt19937_64> streams;  // STREAM mode: one per worker + one engine stream (last)
    Area hospital_initial, hospital_expanded;
    int region = 0;
    std::vector<Update> updates;

    Rng engine_rng(uint32_t agent, uint32_t hour, uint32_t domain) {
        Rng r; r.mode = (mode == Rng::TABLE) ? Rng::KEYED : mode; r.seed = seed; r.agent = agent; r.hour = hour; r.domain = domain;
        if (mode == Rng::STREAM) r.stream = &streams.back();
        return r;
    }

    // Area::random_points (area.rs:64-74): how many points it can return
    static size_t random_points_len(const Area& a, size_t number_of_points) {
        size_t nx = (size_t)std::ceil(std::sqrt((float)number_of_points));
        size_t w = (size_t)(a.end_offset.x - a.start_offset.x + 1), h = (size_t)(a.end_offset.y - a.start_offset.y + 1);
        size_t got = std::min(nx, w) * std::min(nx, h);
        return std::min(got, number_of_points);
    }

    // Epidemiology::new (epidemiology_simulation.rs:75-135) for Population::Auto
    void init(const orc_config& c, uint64_t seed_, Rng::Mode mode_, int threads_) {
        cfg = c; seed = seed_; mode = mode_; threads = std::max(1, threads_);
        cfg.population_csv_file[sizeof(cfg.population_csv_file) - 1] = 0;
        records.clear();
        if (cfg.population_csv_file[0]) {  // Population::Csv (epidemiology_simulation.rs:91)
            records = read_population_records(cfg.population_csv_file);
            cfg.number_of_agents = (uint32_t)records.size();
        }
        disease = disease_from(c);
        streams.clear();
        for (int t = 0; t <= threads; ++t) streams.emplace_back(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull * (uint64_t)(t + 1));
        Grid grid = define_geography(c.grid_size, region);  // :88
        std::vector<Point> start_locations;
        std::vector<Citizen> agent_list;
        generate_population(grid, start_locations, agent_list);  // :92-99
        grid.resize_hospital((int)agent_list.size(), constants::HOSPITAL_STAFF_PERCENTAGE, c.hospital_beds_percentage);  // :101-106
        hospital_initial = grid.hospital_area;
        hospital_expanded = grid.hospital_area;
        hospital_expanded.end_offset = Point{(CoOrdinate)c.grid_size, (CoOrdinate)c.grid_size};
        map.init(grid, agent_list, start_locations);  // :108
        // counts_at_start (utils/util.rs:45-51)
        Count total = c.exposed + c.infected_mild_asymptomatic + c.infected_mild_symptomatic + c.infected_severe;
        counts_at_hr = Counts();
        counts_at_hr.susceptible = map.current_population() - total;
        counts_at_hr.exposed = c.exposed;
        counts_at_hr.infected = total - c.exposed;
        // init_interventions (:178-192)
        interventions = interventions_from(c);
        double ess = interventions.lockdown.get_essential_workers_percentage();
        for (size_t i = 0; i < map.current_locations.capacity(); ++i) {
            if (!map.current_locations.used[i]) continue;
            Citizen& z = map.current_locations.vals[i];
            if (z.work_status == Normal) {  // Citizen::assign_essential_worker citizen/mod.rs:434-440
                Rng r = engine_rng(creation_of[z.id], 0, DOM_INIT);
                if (r.gen_bool(IS_ESSENTIAL, ess)) z.work_status = Essential;
            }
        }
        events.clear();
    }

    // Grid::generate_population (grid.rs:83-123) + citizen_factory (citizen_factory.rs:31-88)
    // + set_start_locations_and_occupancies (grid.rs:125-155)
    // With records (Population::Csv): Grid::read_population (grid.rs:194-231) + Citizen::from_record (citizen/mod.rs:155-180):
    // record c gets houses.cycle()[c] / offices.cycle()[c]; working and uses_public_transport come from the file.
    void generate_population(Grid& grid, std::vector<Point>& home_loc, std::vector<Citizen>& agents) {
        const bool from_csv = cfg.population_csv_file[0] != 0;
        const uint32_t n = cfg.number_of_agents;
        if (from_csv && (size_t)n > (size_t)(constants::HOME_SIZE * constants::HOME_SIZE) * grid.houses.size())
            throw std::runtime_error("Cannot accommodate citizens into homes!");  // grid.rs:216-223
        if (grid.houses.empty() || grid.offices.empty()) throw std::runtime_error("grid too small: no houses or offices");
        double n_pt = (double)n * (cfg.public_transport_percentage + 0.1) * (cfg.working_percentage + 0.1);  // grid.rs:98-99
        size_t n_transport_locations = random_points_len(grid.transport_area, (size_t)std::ceil(n_pt));
        size_t pt_users = 0;
        const size_t H = grid.houses.size(), O = grid.offices.size();
        // The citizens in the reference's creation order: number c gets houses[c % H] and offices[c % O] (grid.rs:108-113,
        // 205-206).  Init draws are keyed on c.
        std::vector<Citizen> made(n);
        for (uint32_t c = 0; c < n; ++c) {  // create_citizen citizen_factory.rs:58-88 / Citizen::from_record citizen/mod.rs:155-180
            Rng r = engine_rng(c, 0, DOM_INIT);
            bool is_working = from_csv ? records[c].working : r.gen_bool(IS_WORKING, cfg.working_percentage);
            Area home = grid.houses[c % H];
            Area work = grid.offices[c % O];
            bool uses_pt = from_csv ? records[c].pub_transport
                                    : r.gen_bool(IS_PT, cfg.public_transport_percentage) && is_working && pt_users < n_transport_locations;
            if (uses_pt) pt_users++;
            Citizen z;
            z.id = c;
            z.home_location = home;
            z.work_location = is_working ? work : home;
            z.transport_location = home.start_offset;  // never read on the hot path (grid.rs:202 TODO in reference)
            z.uses_public_transport = uses_pt;
            // derive_work_status citizen/mod.rs:442-450
            if (is_working) z.work_status = r.gen_bool(IS_STAFF, constants::HOSPITAL_STAFF_PERCENTAGE) ? HospitalStaff : Normal;
            else z.work_status = NA;
            z.work_start_at = constants::ROUTINE_WORK_TIME;
            z.immunity = constants::IMMUNITY_RANGE[r.choose_index(IS_IMMUNITY, 5)];  // :210-213
            z.current_area = home;
            made[c] = z;
        }
        // start locations: Area::random_points(k) inside each agent's own house (area.rs:64-74).
        // k>=2 -> both xs and both ys are chosen (order preserved), x outer / y inner, take k; k==1 -> one random x, one random y.
        std::vector<uint8_t> per_house(H, 0);
        for (uint32_t c = 0; c < n; ++c) {
            if (per_house[c % H] >= constants::HOME_SIZE * constants::HOME_SIZE)
                throw std::runtime_error("There are more agents assigned to a house than house capacity");  // grid.rs:140-142
            per_house[c % H]++;
        }
        std::vector<Point> made_loc(n);
        for (uint32_t c = 0; c < n; ++c) {
            const Area& home = made[c].home_location;
            const uint32_t k = per_house[c % H], j = c / (uint32_t)H;
            if (k == 1) {
                Rng r = engine_rng(c, 0, DOM_INIT);
                int x = home.start_offset.x + (int)r.choose_index(IS_STARTX, 2);
                int y = home.start_offset.y + (int)r.choose_index(IS_STARTY, 2);
                made_loc[c] = Point{x, y};
            } else {
                made_loc[c] = Point{home.start_offset.x + (int)(j / 2), home.start_offset.y + (int)(j % 2)};
            }
        }
        // Agent numbering -- a convention shared with the engine under test, not reference behaviour (the reference's ids
        // are opaque Uuids): ids follow (house, creation number), i.e. the population is numbered house by house.
        // EPI_AGENT_ORDER=creation keeps id == creation number.
        const char* order_env = getenv("EPI_AGENT_ORDER");
        creation_of.resize(n);
        for (uint32_t c = 0; c < n; ++c) creation_of[c] = c;
        if (!(order_env && std::string(order_env) == "creation"))
            std::stable_sort(creation_of.begin(), creation_of.end(),
                             [&](uint32_t x, uint32_t y) { return std::make_tuple(x % (uint32_t)H, x / (uint32_t)H) < std::make_tuple(y % (uint32_t)H, y / (uint32_t)H); });
        agents.resize(n);
        home_loc.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            agents[i] = made[creation_of[i]];
            agents[i].id = i;
            home_loc[i] = made_loc[creation_of[i]];
        }
        // set_starting_infections citizen_factory.rs:112-134 : choose_multiple (uniform, without replacement)
        Count total = cfg.exposed + cfg.infected_mild_asymptomatic + cfg.infected_mild_symptomatic + cfg.infected_severe;
        if (total > n) throw std::runtime_error("more starting infections than agents");
        std::vector<uint32_t> chosen;
        std::unordered_set<uint32_t> seen;
        for (uint32_t k = 0; chosen.size() < total; ++k) {
            Rng r = engine_rng(k, 0, DOM_STARTINF);
            uint32_t idx = r.choose_index(0, n);
            if (seen.insert(idx).second) chosen.push_back(idx);
        }
        size_t q = 0;
        for (Count j = 0; j < cfg.exposed; ++j) agents[chosen[q++]].state_machine.state = State::expose(0);
        for (Count j = 0; j < cfg.infected_mild_asymptomatic; ++j) agents[chosen[q++]].state_machine.state = State::infected(1, Asymptomatic);
        for (Count j = 0; j < cfg.infected_mild_symptomatic; ++j) agents[chosen[q++]].state_machine.state = State::infected(1, Mild);
        for (Count j = 0; j < cfg.infected_severe; ++j) agents[chosen[q++]].state_machine.state = State::infected(1, Severe);
    }

    // CitizenLocationMap::simulate (allocation_map.rs:67-129), standalone (travel_plan_config == None)
    void simulate(Counts& csv_record, Hour simulation_hour, const uint64_t* draw_table) {
        csv_record.clear();
        if (simulation_hour % 24 == constants::ROUTINE_START_TIME) map.prime_hospital_cache(); else map.hospital_cache_valid = false;
        const size_t cap = map.current_locations.capacity();
        updates.clear();
        // PHASE A (:82-92): rayon par_iter over the start-of-hour map
        std::vector<std::vector<Update>> per_thread((size_t)threads);
#pragma omp parallel num_threads(threads)
        {
            int t = omp_get_thread_num();
            std::vector<Update>& out = per_thread[(size_t)t];
            out.reserve(map.current_locations.len() / (size_t)threads + 64);
            Rng rng; rng.seed = seed; rng.hour = simulation_hour; rng.domain = DOM_STEP;
            rng.mode = draw_table ? Rng::TABLE : mode;
            if (mode == Rng::STREAM) rng.stream = &streams[(size_t)t];
#pragma omp for schedule(static)
            for (long s = 0; s < (long)cap; ++s) {
                if (!map.current_locations.used[(size_t)s]) continue;
                Update u;
                u.old_cell = map.current_locations.keys[(size_t)s];
                u.agent = map.current_locations.vals[(size_t)s];  // let mut current_agent = *agent
                rng.agent = u.agent.id;
                if (draw_table) rng.row = draw_table + (size_t)u.agent.id * SLOTS_PER_AGENT;
                u.infection_status = u.agent.state_machine.is_infected();
                u.new_cell = u.agent.routine(u.old_cell, simulation_hour, map.grid, map, rng, disease);
                out.push_back(u);
            }
        }
        for (auto& v : per_thread) updates.insert(updates.end(), v.begin(), v.end());
        // phase-B order: the reference walks `updates` in hash-map iteration order (arbitrary).  KEYED/TABLE
        // modes use ascending agent id -- the GPU's lowest-id-wins rule.
        if (mode != Rng::STREAM || draw_table)
            std::sort(updates.begin(), updates.end(), [](const Update& a, const Update& b) { return a.agent.id < b.agent.id; });
        else if (shuffle_phase_b)
            std::shuffle(updates.begin(), updates.end(), streams.back());
        // PHASE B (:93-125)
        for (const Update& u : updates) {
            const Citizen& agent_at_new_cell = map.upcoming_locations.entry_or_insert(u.new_cell, u.agent);
            if (agent_at_new_cell.id != u.agent.id) map.upcoming_locations.insert(u.old_cell, u.agent);
            csv_record.update_counts(u.agent);
        }
        map.swap();  // :127
        if (csv_record.total() != map.current_population()) throw std::runtime_error("assert_eq!(csv_record.total(), current_population) failed");  // :128
    }

    // CitizenLocationMap::vaccinate (allocation_map.rs:381-387)
    void vaccinate(double p, Hour hour) {
        for (size_t i = 0; i < map.current_locations.capacity(); ++i) {
            if (!map.current_locations.used[i]) continue;
            Citizen& z = map.current_locations.vals[i];
            if (z.state_machine.is_susceptible()) {
                Rng r = engine_rng(z.id, hour, DOM_VACCINATE);
                if (r.gen_bool(0, p)) z.vaccinated = true;
            }
        }
    }
    void expand_hospital() { map.grid.increase_hospital_size(cfg.grid_size); }

    // CitizenLocationMap::process_interventions (allocation_map.rs:306-337)
    void process_interventions() {
        const Counts& c = counts_at_hr;
        double p;
        if (interventions.vaccinate.get_vaccination_percentage(c, p)) {  // apply_vaccination_intervention :367-379
            vaccinate(p, c.hour);
            events.push_back({c.hour, 1, 0});
        }
        if (interventions.lockdown.should_apply(c)) {
            interventions.lockdown.apply();
            map.lock_city();
            events.push_back({c.hour, 0, 1});
        }
        if (interventions.lockdown.should_unlock(c)) {
            map.unlock_city();
            interventions.lockdown.unapply();
            events.push_back({c.hour, 0, 0});
        }
        interventions.build_new_hospital.counts_updated(c);
        if (interventions.build_new_hospital.should_apply(c)) {
            expand_hospital();
            interventions.build_new_hospital.apply();
            events.push_back({c.hour, 2, 0});
        }
    }

    // Epidemiology::stop_simulation (epidemiology_simulation.rs:564-575), Standalone arm
    static bool stop_simulation(const Counts& row) { return row.exposed == 0 && row.infected == 0 && row.hospitalized == 0; }

    // Epidemiology::run_single_engine (epidemiology_simulation.rs:211-274); rows = what CsvListener collects
    size_t run_single_engine(std::vector<Counts>& rows, Hour max_hours = 0) {
        Hour hours = max_hours ? std::min(max_hours, cfg.hours) : cfg.hours;
        for (Hour simulation_hour = 1; simulation_hour < hours; ++simulation_hour) {
            counts_at_hr.increment_hour();
            if (map.current_population() == 0) throw std::runtime_error("No citizens!");
            simulate(counts_at_hr, simulation_hour, nullptr);
            rows.push_back(counts_at_hr);
            process_interventions();
            if (stop_simulation(counts_at_hr)) break;
        }
        return rows.size();
    }
};

// -----------------------------------------------------------------------------
// Packed per-agent words: the exchange format between oracle and GPU engine in the parity tests.
// Normative layout: DESIGN.md "Agent state word"; restated independently in epirust_b200/csrc/layout.h.
// -----------------------------------------------------------------------------
enum AreaKind : uint32_t { AK_HOME = 0, AK_WORK = 1, AK_TRANSPORT = 2, AK_HOUSING = 3, AK_HOSPITAL0 = 4, AK_HOSPITAL1 = 5 };

inline uint32_t house_index(const Grid& g, const Area& a) {
    int nx = (g.housing_area.end_offset.x - g.housing_area.start_offset.x + 1) / (int)constants::HOME_SIZE;
    return (uint32_t)(((a.start_offset.y - g.housing_area.start_offset.y) / (int)constants::HOME_SIZE) * nx +
                      (a.start_offset.x - g.housing_area.start_offset.x) / (int)constants::HOME_SIZE);
}
inline uint32_t office_index(const Grid& g, const Area& a) {
    int nx = (g.work_area.end_offset.x - g.work_area.start_offset.x + 1) / (int)constants::OFFICE_SIZE;
    return (uint32_t)(((a.start_offset.y - g.work_area.start_offset.y) / (int)constants::OFFICE_SIZE) * nx +
                      (a.start_offset.x - g.work_area.start_offset.x) / (int)constants::OFFICE_SIZE);
}

inline uint32_t pack_state_word(const Engine& e, const Citizen& z) {
    const Grid& g = e.map.grid;
    uint32_t kind;
    if (z.current_area == z.home_location) kind = AK_HOME;
    else if (z.current_area == z.work_location) kind = AK_WORK;
    else if (z.current_area == g.transport_area) kind = AK_TRANSPORT;
    else if (z.current_area == g.housing_area) kind = AK_HOUSING;
    else if (z.current_area == e.hospital_initial) kind = AK_HOSPITAL0;
    else if (z.current_area == e.hospital_expanded) kind = AK_HOSPITAL1;
    else throw std::runtime_error("current_area is not a known area");
    const State& s = z.state_machine.state;
    uint32_t w = 0;
    w |= (uint32_t)s.kind;
    w |= (uint32_t)(s.kind == Infected ? s.severity : 0) << 3;
    w |= (uint32_t)(z.immunity + 2) << 5;
    w |= (uint32_t)z.vaccinated << 8;
    w |= (uint32_t)z.uses_public_transport << 9;
    w |= (uint32_t)z.hospitalized << 10;
    w |= (uint32_t)z.isolated << 11;
    w |= (uint32_t)z.work_quarantined << 12;
    w |= (uint32_t)z.work_status << 13;
    w |= kind << 15;
    w |= (uint32_t)(s.kind == Infected ? (s.infection_day & 0x3FFFu) : 0) << 18;
    return w;
}

inline Citizen unpack_citizen(const Engine& e, uint32_t id, uint32_t w, uint32_t t0, uint32_t home, uint32_t work, uint32_t wsa) {
    const Grid& g = e.map.grid;
    Citizen z;
    z.id = id;
    State s;
    s.kind = (StateKind)(w & 7u);
    s.severity = (SeverityKind)((w >> 3) & 3u);
    s.infection_day = (w >> 18) & 0x3FFFu;
    s.at_hour = t0;
    if (s.kind != Infected) { s.severity = Pre; s.infection_day = 0; }
    if (!(s.kind == Exposed || (s.kind == Infected && s.severity == Pre))) s.at_hour = 0;
    z.state_machine.state = s;
    z.immunity = (int)((w >> 5) & 7u) - 2;
    z.vaccinated = (w >> 8) & 1u;
    z.uses_public_transport = (w >> 9) & 1u;
    z.hospitalized = (w >> 10) & 1u;
    z.isolated = (w >> 11) & 1u;
    z.work_quarantined = (w >> 12) & 1u;
    z.work_status = (WorkStatusKind)((w >> 13) & 3u);
    z.home_location = g.houses.at(home);
    z.work_location = z.work_status == NA ? z.home_location : g.offices.at(work);
    z.work_start_at = wsa;
    z.transport_location = z.home_location.start_offset;
    switch ((w >> 15) & 7u) {
        case AK_HOME: z.current_area = z.home_location; break;
        case AK_WORK: z.current_area = z.work_location; break;
        case AK_TRANSPORT: z.current_area = g.transport_area; break;
        case AK_HOUSING: z.current_area = g.housing_area; break;
        case AK_HOSPITAL0: z.current_area = e.hospital_initial; break;
        case AK_HOSPITAL1: z.current_area = e.hospital_expanded; break;
        default: throw std::runtime_error("bad area kind");
    }
    return z;
}

}  // namespace orc
