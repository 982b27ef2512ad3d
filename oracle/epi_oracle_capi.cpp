// =============================================================================
// oracle/epi_oracle_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench.py's
// cpu_baseline and --impl reference legs).  See epi_oracle.hpp for parity status.
// =============================================================================
#include <chrono>
#include <cstdio>

#include "epi_oracle_engine.hpp"

using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret)                  \
    }                                   \
    catch (const std::exception& ex) {  \
        g_err = ex.what();              \
        return ret;                     \
    }

static void counts_out(const Counts& c, uint32_t* out) {
    out[0] = c.hour; out[1] = c.susceptible; out[2] = c.exposed; out[3] = c.infected;
    out[4] = c.hospitalized; out[5] = c.recovered; out[6] = c.deceased;
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- engine -------------------------------------------------------------------
// mode: 0 KEYED (Philox, id-ordered phase B), 2 STREAM (sequential mt19937_64, hash-order phase B)
void* orc_create(const orc_config* cfg, uint64_t seed, int mode, int threads) {
    ORC_TRY
    Engine* e = new Engine();
    e->init(*cfg, seed, mode == 2 ? Rng::STREAM : Rng::KEYED, threads);
    return e;
    ORC_CATCH(nullptr)
}
void orc_destroy(void* h) { delete (Engine*)h; }
uint32_t orc_population(void* h) { return ((Engine*)h)->map.current_population(); }
void orc_set_shuffle_phase_b(void* h, int on) { ((Engine*)h)->shuffle_phase_b = on != 0; }

void orc_counts_at_start(void* h, uint32_t* out7) { counts_out(((Engine*)h)->counts_at_hr, out7); }

int orc_step(void* h, uint32_t hour, uint32_t* out7) {
    ORC_TRY
    Engine* e = (Engine*)h;
    e->counts_at_hr.hour = hour;
    e->simulate(e->counts_at_hr, hour, nullptr);
    counts_out(e->counts_at_hr, out7);
    return 0;
    ORC_CATCH(1)
}
// draws: [population][16] u64, row = agent id
int orc_step_with_draws(void* h, uint32_t hour, const uint64_t* draws, uint32_t* out7) {
    ORC_TRY
    Engine* e = (Engine*)h;
    e->counts_at_hr.hour = hour;
    e->simulate(e->counts_at_hr, hour, draws);
    counts_out(e->counts_at_hr, out7);
    return 0;
    ORC_CATCH(1)
}
int orc_lock_city(void* h) { ((Engine*)h)->map.lock_city(); return 0; }
int orc_unlock_city(void* h) { ((Engine*)h)->map.unlock_city(); return 0; }
int orc_vaccinate(void* h, double p, uint32_t hour) { ((Engine*)h)->vaccinate(p, hour); return 0; }
int orc_expand_hospital(void* h) { ((Engine*)h)->expand_hospital(); return 0; }

// geometry: out[0..3] housing sx,sy,ex,ey ; [4..7] transport ; [8..11] work ; [12..15] hospital (current)
//           [16] n_houses [17] n_offices [18] grid_size
void orc_geometry(void* h, int32_t* out) {
    const Grid& g = ((Engine*)h)->map.grid;
    const Area* as[4] = {&g.housing_area, &g.transport_area, &g.work_area, &g.hospital_area};
    for (int i = 0; i < 4; ++i) {
        out[4 * i + 0] = as[i]->start_offset.x; out[4 * i + 1] = as[i]->start_offset.y;
        out[4 * i + 2] = as[i]->end_offset.x;   out[4 * i + 3] = as[i]->end_offset.y;
    }
    out[16] = (int32_t)g.houses.size(); out[17] = (int32_t)g.offices.size(); out[18] = (int32_t)g.grid_size;
}

// state in agent-id order: cell_x, cell_y, st (packed), t0, home, work, wsa  (arrays of length population)
int orc_get_state(void* h, int32_t* cx, int32_t* cy, uint32_t* st, uint32_t* t0, uint32_t* home, uint32_t* work, uint32_t* wsa) {
    ORC_TRY
    Engine* e = (Engine*)h;
    const PointMap& m = e->map.current_locations;
    const uint32_t n = (uint32_t)m.len();
    for (size_t i = 0; i < m.capacity(); ++i) {
        if (!m.used[i]) continue;
        const Citizen& z = m.vals[i];
        if (z.id >= n) throw std::runtime_error("agent id out of range");
        cx[z.id] = m.keys[i].x; cy[z.id] = m.keys[i].y;
        st[z.id] = pack_state_word(*e, z);
        const State& s = z.state_machine.state;
        t0[z.id] = (s.kind == Exposed || (s.kind == Infected && s.severity == Pre)) ? s.at_hour : 0;
        home[z.id] = house_index(e->map.grid, z.home_location);
        work[z.id] = z.work_status == NA ? 0 : office_index(e->map.grid, z.work_location);
        wsa[z.id] = z.work_status == HospitalStaff ? z.work_start_at : 0;
    }
    return 0;
    ORC_CATCH(1)
}
int orc_set_state(void* h, uint32_t n, const int32_t* cx, const int32_t* cy, const uint32_t* st, const uint32_t* t0, const uint32_t* home,
                  const uint32_t* work, const uint32_t* wsa) {
    ORC_TRY
    Engine* e = (Engine*)h;
    e->map.current_locations.init(n);
    e->map.upcoming_locations.init(n);
    for (uint32_t i = 0; i < n; ++i) {
        Citizen z = unpack_citizen(*e, i, st[i], t0[i], home[i], work[i], wsa[i]);
        if (e->map.current_locations.insert(Point{cx[i], cy[i]}, z)) throw std::runtime_error("two agents on one cell");
    }
    return 0;
    ORC_CATCH(1)
}

// Whole standalone run (run_single_engine).  rows: [max_rows][7]; events: [max_events][3] (hour, kind, status).
// Returns number of rows, or -1.  *n_events receives the number of intervention events, *seconds the hour-loop wall time.
long orc_run(const orc_config* cfg, uint64_t seed, int mode, int threads, uint32_t max_hours, uint32_t* rows, long max_rows, uint32_t* events,
             long max_events, long* n_events, double* seconds) {
    ORC_TRY
    Engine e;
    e.init(*cfg, seed, mode == 2 ? Rng::STREAM : Rng::KEYED, threads);
    std::vector<Counts> r;
    auto t0 = std::chrono::steady_clock::now();
    e.run_single_engine(r, max_hours);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    long n = std::min<long>((long)r.size(), max_rows);
    for (long i = 0; i < n; ++i) counts_out(r[(size_t)i], rows + 7 * i);
    long ne = std::min<long>((long)e.events.size(), max_events);
    for (long i = 0; i < ne; ++i) { events[3 * i] = e.events[(size_t)i].hour; events[3 * i + 1] = (uint32_t)e.events[(size_t)i].kind; events[3 * i + 2] = (uint32_t)e.events[(size_t)i].status; }
    if (n_events) *n_events = ne;
    return (long)r.size();
    ORC_CATCH(-1)
}

// Timed sample for the CPU baseline: runs hours [first_hour, first_hour+n_hours) on an existing engine, returns seconds.
double orc_time_hours(void* h, uint32_t first_hour, uint32_t n_hours) {
    Engine* e = (Engine*)h;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t hr = first_hour; hr < first_hour + n_hours; ++hr) {
        e->counts_at_hr.hour = hr;
        e->simulate(e->counts_at_hr, hr, nullptr);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- known-answer helpers (tests/test_oracle_kat.py) ----------------------------
void orc_kat_philox(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { philox4x32_10(ctr4, key2, out4); }
uint64_t orc_kat_draw(uint64_t seed, uint32_t agent, uint32_t hour, uint32_t domain, uint32_t slot) {
    Rng r; r.mode = Rng::KEYED; r.seed = seed; r.agent = agent; r.hour = hour; r.domain = domain;
    return r.next(slot);
}
uint32_t orc_kat_draw32(uint64_t seed, uint32_t agent, uint32_t hour, uint32_t slot) {
    Rng r; r.mode = Rng::KEYED; r.seed = seed; r.agent = agent; r.hour = hour; r.domain = DOM_STEP;
    return r.next32(slot);
}
uint64_t orc_kat_bernoulli_threshold(double p) { return bernoulli_threshold(p); }
void orc_kat_neighbors(int x, int y, int32_t* out16) {
    for (int j = 0; j < 8; ++j) { out16[2 * j] = x + NEIGHBOR_OFFSETS[j][0]; out16[2 * j + 1] = y + NEIGHBOR_OFFSETS[j][1]; }
}
static Area mk_area(int sx, int sy, int ex, int ey) { Area a; a.start_offset = Point{sx, sy}; a.end_offset = Point{ex, ey}; return a; }
int orc_kat_area_neighbors(int sx, int sy, int ex, int ey, int px, int py, int32_t* out16) {
    Point nb[8];
    int n = mk_area(sx, sy, ex, ey).get_neighbors_of(Point{px, py}, nb);
    for (int j = 0; j < n; ++j) { out16[2 * j] = nb[j].x; out16[2 * j + 1] = nb[j].y; }
    return n;
}
int orc_kat_area_contains(int sx, int sy, int ex, int ey, int px, int py) { return mk_area(sx, sy, ex, ey).contains(Point{px, py}); }
uint32_t orc_kat_number_of_cells(int sx, int sy, int ex, int ey) { return mk_area(sx, sy, ex, ey).get_number_of_cells(); }
long orc_kat_area_iter(int sx, int sy, int ex, int ey, int32_t* out, long max_points) {
    auto v = mk_area(sx, sy, ex, ey).iter_all();
    long n = std::min<long>((long)v.size(), max_points);
    for (long i = 0; i < n; ++i) { out[2 * i] = v[(size_t)i].x; out[2 * i + 1] = v[(size_t)i].y; }
    return (long)v.size();
}
long orc_kat_area_factory(int sx, int sy, int ex, int ey, uint32_t size, int32_t* out, long max_areas) {
    auto v = area_factory(Point{sx, sy}, Point{ex, ey}, size, 0);
    long n = std::min<long>((long)v.size(), max_areas);
    for (long i = 0; i < n; ++i) {
        out[4 * i] = v[(size_t)i].start_offset.x; out[4 * i + 1] = v[(size_t)i].start_offset.y;
        out[4 * i + 2] = v[(size_t)i].end_offset.x; out[4 * i + 3] = v[(size_t)i].end_offset.y;
    }
    return (long)v.size();
}
static void grid_out(const Grid& g, int32_t* out) {
    const Area* as[4] = {&g.housing_area, &g.transport_area, &g.work_area, &g.hospital_area};
    for (int i = 0; i < 4; ++i) {
        out[4 * i + 0] = as[i]->start_offset.x; out[4 * i + 1] = as[i]->start_offset.y;
        out[4 * i + 2] = as[i]->end_offset.x;   out[4 * i + 3] = as[i]->end_offset.y;
    }
    out[16] = (int32_t)g.houses.size(); out[17] = (int32_t)g.offices.size(); out[18] = (int32_t)g.grid_size;
}
void orc_kat_define_geography(uint32_t grid_size, int32_t* out19) { grid_out(define_geography(grid_size, 0), out19); }
void orc_kat_resize_hospital(uint32_t grid_size, int n_agents, double staff_pct, double beds_pct, int32_t* out19) {
    Grid g = define_geography(grid_size, 0);
    g.resize_hospital(n_agents, staff_pct, beds_pct);
    grid_out(g, out19);
}
void orc_kat_increase_hospital(uint32_t grid_size, uint32_t new_size, int32_t* out19) {
    Grid g = define_geography(grid_size, 0);
    g.increase_hospital_size(new_size);
    grid_out(g, out19);
}
double orc_kat_transmission_rate(const orc_config* cfg, uint32_t infection_day) { return disease_from(*cfg).get_current_transmission_rate(infection_day); }
int orc_kat_is_to_be_hospitalized(const orc_config* cfg, uint32_t infection_day) { return disease_from(*cfg).is_to_be_hospitalized(infection_day); }

// goto_hospital scenario (allocation_map.rs:445-495): 2 agents at given points on a grid of `grid_size`, hospital rect given;
// returns is_hospitalized and the new location of agent 0 whose home is (hsx,hsy)-(hex,hey).
int orc_kat_goto_hospital(uint32_t grid_size, const int32_t* occupied_xy, int n_occupied, int hosp_sx, int hosp_sy, int hosp_ex, int hosp_ey,
                          int home_sx, int home_sy, int home_ex, int home_ey, int cell_x, int cell_y, uint64_t seed, int32_t* out_xy) {
    CitizenLocationMap map;
    Grid g = define_geography(grid_size, 0);
    std::vector<Citizen> agents((size_t)n_occupied);
    std::vector<Point> pts((size_t)n_occupied);
    for (int i = 0; i < n_occupied; ++i) { agents[(size_t)i].id = (uint32_t)i; pts[(size_t)i] = Point{occupied_xy[2 * i], occupied_xy[2 * i + 1]}; }
    map.init(g, agents, pts);
    Citizen z; z.id = 0; z.home_location = mk_area(home_sx, home_sy, home_ex, home_ey);
    Rng r; r.mode = Rng::KEYED; r.seed = seed;
    auto res = map.goto_hospital(mk_area(hosp_sx, hosp_sy, hosp_ex, hosp_ey), Point{cell_x, cell_y}, z, r);
    out_xy[0] = res.second.x; out_xy[1] = res.second.y;
    return res.first ? 1 : 0;
}
int orc_kat_is_point_in_grid(uint32_t grid_size, int x, int y) {
    CitizenLocationMap map; map.grid = define_geography(grid_size, 0);
    return map.is_point_in_grid(Point{x, y});
}

// interventions (lockdown.rs / hospital.rs / vaccination.rs tests).  A tiny scriptable harness:
// op 0: lockdown.should_apply(counts) ; 1: lockdown.apply ; 2: should_unlock ; 3: unapply ; 4: set_zero_infection_hour(arg)
// op 5: hospital.counts_updated ; 6: hospital.should_apply ; 7: hospital.apply ; 8: vaccination percentage*1e6 or -1
void* orc_iv_create(const orc_config* cfg) { return new Interventions(interventions_from(*cfg)); }
void orc_iv_destroy(void* p) { delete (Interventions*)p; }
long orc_iv_op(void* p, int op, const uint32_t* c7, uint32_t arg) {
    Interventions& iv = *(Interventions*)p;
    Counts c; c.hour = c7[0]; c.susceptible = c7[1]; c.exposed = c7[2]; c.infected = c7[3]; c.hospitalized = c7[4]; c.recovered = c7[5]; c.deceased = c7[6];
    switch (op) {
        case 0: return iv.lockdown.should_apply(c);
        case 1: iv.lockdown.apply(); return iv.lockdown.is_locked_down;
        case 2: return iv.lockdown.should_unlock(c);
        case 3: iv.lockdown.unapply(); return iv.lockdown.is_locked_down;
        case 4: iv.lockdown.set_zero_infection_hour(arg); return iv.lockdown.zero_infection_hour;
        case 5: iv.build_new_hospital.counts_updated(c); return iv.build_new_hospital.new_infections_in_a_day;
        case 6: return iv.build_new_hospital.should_apply(c);
        case 7: iv.build_new_hospital.apply(); return iv.build_new_hospital.has_applied;
        case 8: { double pc; return iv.vaccinate.get_vaccination_percentage(c, pc) ? (long)std::llround(pc * 1e6) : -1; }
    }
    return -2;
}
// Counts::update_counts over a packed-state list (counts.rs:126-140)
void orc_kat_counts(const uint32_t* st, uint32_t n, uint32_t* out7) {
    Counts c;
    for (uint32_t i = 0; i < n; ++i) {
        Citizen z; z.state_machine.state.kind = (StateKind)(st[i] & 7u); z.hospitalized = (st[i] >> 10) & 1u;
        c.update_counts(z);
    }
    counts_out(c, out7);
}

}  // extern "C"

// ---- multi-region (oracle/epi_oracle_travel.hpp) -------------------------------------------------------------------
#include "epi_oracle_travel.hpp"

extern "C" {

void* orc_multi_create(const orc_config* cfgs, int n_regions, uint64_t seed, const uint32_t* migration, const uint32_t* commute, int migration_enabled,
                       int commute_enabled, uint32_t start_migration_hour, uint32_t end_migration_hour, uint32_t extra_capacity, int threads) {
    ORC_TRY
    TravelPlanConfig p;
    p.n_regions = n_regions;
    p.migration_enabled = migration_enabled != 0;
    p.commute_enabled = commute_enabled != 0;
    p.migration.assign(migration, migration + (size_t)n_regions * n_regions);
    p.commute.assign(commute, commute + (size_t)n_regions * n_regions);
    p.start_migration_hour = start_migration_hour;
    p.end_migration_hour = end_migration_hour;
    MultiEngine* m = new MultiEngine();
    m->init(std::vector<orc_config>(cfgs, cfgs + n_regions), seed, p, extra_capacity, threads);
    return m;
    ORC_CATCH(nullptr)
}
void orc_multi_destroy(void* h) { delete (MultiEngine*)h; }
// rows: [n_regions][7]
int orc_multi_step(void* h, uint32_t hour, uint32_t* rows) {
    ORC_TRY
    MultiEngine* m = (MultiEngine*)h;
    m->step(hour);
    for (size_t r = 0; r < m->regions.size(); ++r) counts_out(m->regions[r].counts_at_hr, rows + 7 * r);
    return 0;
    ORC_CATCH(1)
}
uint32_t orc_multi_capacity(void* h, int r) { return ((MultiEngine*)h)->regions[(size_t)r].capacity; }
uint32_t orc_multi_population(void* h, int r) { return ((MultiEngine*)h)->regions[(size_t)r].map.current_population(); }
uint32_t orc_multi_max_place_rounds(void* h, int r) { return ((MultiEngine*)h)->regions[(size_t)r].max_place_rounds; }
// state by slot (arrays of length capacity); absent slots read st = 7, everything else 0.  reg = home region | work region << 8
int orc_multi_get_state(void* h, int r, int32_t* cx, int32_t* cy, uint32_t* st, uint32_t* t0, uint32_t* home, uint32_t* work, uint32_t* wsa, uint32_t* reg) {
    ORC_TRY
    RegionEngine& e = ((MultiEngine*)h)->regions[(size_t)r];
    for (uint32_t s = 0; s < e.capacity; ++s) { cx[s] = cy[s] = 0; st[s] = 7; t0[s] = home[s] = work[s] = wsa[s] = reg[s] = 0; }
    const PointMap& m = e.map.current_locations;
    for (size_t i = 0; i < m.capacity(); ++i) {
        if (!m.used[i]) continue;
        const Citizen& z = m.vals[i];
        if (z.id >= e.capacity) throw std::runtime_error("agent slot out of range");
        cx[z.id] = m.keys[i].x; cy[z.id] = m.keys[i].y;
        st[z.id] = pack_state_word(e, z);
        const State& s = z.state_machine.state;
        t0[z.id] = (s.kind == Exposed || (s.kind == Infected && s.severity == Pre)) ? s.at_hour : 0;
        home[z.id] = house_index(e.map.grid, z.home_location);
        work[z.id] = z.work_status == NA ? 0 : office_index(e.map.grid, z.work_location);
        wsa[z.id] = z.work_status == HospitalStaff ? z.work_start_at : 0;
        reg[z.id] = (uint32_t)z.home_location.location_id | ((uint32_t)z.work_location.location_id << 8);
    }
    return 0;
    ORC_CATCH(1)
}
int orc_multi_events(void* h, int r, uint32_t* events, int max_events) {
    RegionEngine& e = ((MultiEngine*)h)->regions[(size_t)r];
    int n = std::min<int>((int)e.events.size(), max_events);
    for (int i = 0; i < n; ++i) { events[3 * i] = e.events[(size_t)i].hour; events[3 * i + 1] = (uint32_t)e.events[(size_t)i].kind; events[3 * i + 2] = (uint32_t)e.events[(size_t)i].status; }
    return (int)e.events.size();
}

// engine_migration_plan.rs:120-150 known answers: percent_outgoing and the proportional allocation
double orc_kat_percent_outgoing(const uint32_t* matrix, int n_regions, int region, uint32_t population) {
    TravelPlanConfig p; p.n_regions = n_regions; p.migration.assign(matrix, matrix + (size_t)n_regions * n_regions);
    return (double)p.get_total_outgoing(p.migration, region) / (double)population;
}
void orc_kat_alloc_outgoing(const uint32_t* matrix, int n_regions, int region, uint32_t total, uint32_t* counts_out_per_region) {
    TravelPlanConfig p; p.n_regions = n_regions; p.migration.assign(matrix, matrix + (size_t)n_regions * n_regions);
    const uint32_t planned_total = p.get_total_outgoing(p.migration, region);
    uint32_t front = 0;
    for (int to = 0; to < n_regions; ++to) {
        counts_out_per_region[to] = 0;
        if (to == region || p.get_outgoing(p.migration, region, to) == 0) continue;
        const double share = (double)p.get_outgoing(p.migration, region, to) / (double)planned_total;
        uint32_t count = (uint32_t)(int32_t)(share * (double)(int32_t)total);
        if (count > total - front) count = total - front;
        counts_out_per_region[to] = count;
        front += count;
    }
}

}  // extern "C"
