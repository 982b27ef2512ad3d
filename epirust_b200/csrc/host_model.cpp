#include "host_model.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <string>
#include <stdexcept>
#include <unordered_set>

#include "philox.cuh"

namespace epi {

namespace {
// init draw slots (domain DOM_INIT, hour word 0); see DESIGN.md "Draw slots"
enum : uint32_t { IS_WORKING = 0, IS_PT = 1, IS_STAFF = 2, IS_IMMUNITY = 3, IS_ESSENTIAL = 4, IS_STARTX = 5, IS_STARTY = 6 };
constexpr double kHospitalStaffPercentage = 0.002;  // models/constants.rs:43
constexpr uint32_t kRoutineWorkTime = 8;            // models/constants.rs:32

inline int ceil_frac(uint32_t g, double f) { return (int)std::ceil((double)g * f); }
}  // namespace

// Agent numbering.  The reference hands houses and offices out round-robin in creation order (houses[c % H], offices[c % O],
// grid.rs:108-113), so the citizen created as number c shares house c % H with c+H, c+2H, ...; agent identities themselves
// are arbitrary (Uuid v4).  We build exactly that population and then NUMBER it house by house: housemates sit in adjacent
// slots of every per-agent array, so the claim words and grid bytes of a house are touched by neighbouring threads of one
// kernel instead of threads millions of agents apart (each 32-byte claim sector makes one trip to DRAM per kernel instead
// of two at the home hours: -5 % time per simulated day at 10 M agents).  EPI_AGENT_ORDER=creation keeps id == creation
// number (the A/B switch behind that figure); the oracle follows the same variable.  Measured and rejected: numbering the
// public-transport commuters last, so that the warps of the home hours are uniform -- the random accesses of the transport
// strip then pile up at the end of every kernel instead of being spread over it (+3 % time).
AgentOrder agent_order() {
    static const AgentOrder v = [] {
        const char* e = getenv("EPI_AGENT_ORDER");
        return e && std::string(e) == "creation" ? AgentOrder::Creation : AgentOrder::House;
    }();
    return v;
}

Geometry make_geometry(uint32_t grid_size, uint32_t n_agents, double beds_pct) {
    // Vertical strips, 40 % housing / 20 % transport / 20 % work / 10 % hospital of the width, each G+1 rows tall
    // because Area ends are inclusive and the reference passes grid_size as the end row (geography/mod.rs:43-50).
    Geometry g;
    g.grid_size = (int)grid_size;
    const int G = (int)grid_size;
    int x0 = 0;
    auto strip = [&](double frac) {
        const int w = ceil_frac(grid_size, frac);
        Rect r{x0, 0, x0 + w - 1, G};
        x0 += w;
        return r;
    };
    g.housing = strip(0.4);
    g.transport = strip(0.2);
    g.work = strip(0.2);
    g.hospital_initial = strip(0.1);
    // area_factory: whole tiles only (geography/area.rs:95-117)
    g.house_nx = (g.housing.ex - g.housing.sx + 1) / 2;
    g.house_ny = (g.housing.ey - g.housing.sy + 1) / 2;
    g.office_nx = (g.work.ex - g.work.sx + 1) / 10;
    g.office_ny = (g.work.ey - g.work.sy + 1) / 10;
    g.n_houses = (uint32_t)std::max(0, g.house_nx) * (uint32_t)std::max(0, g.house_ny);
    g.n_offices = (uint32_t)std::max(0, g.office_nx) * (uint32_t)std::max(0, g.office_ny);
    // Grid::resize_hospital (grid.rs:240-261): beds = ceil(N*beds% + N*staff%); shrink the strip to beds / (width-1) rows
    // when the bed count fits in Area::get_number_of_cells (which is (ex-sx)*(ey-sy), area.rs:90-92)
    g.hospital_resized = g.hospital_initial;
    const uint32_t beds = (uint32_t)std::ceil((double)n_agents * beds_pct + (double)n_agents * kHospitalStaffPercentage);
    const uint32_t dx = (uint32_t)(g.hospital_initial.ex - g.hospital_initial.sx);
    const uint32_t cells = (uint32_t)((g.hospital_initial.ex - g.hospital_initial.sx) * (g.hospital_initial.ey - g.hospital_initial.sy));
    if (beds <= cells && dx > 0) g.hospital_resized.ey = (int)(beds / dx);
    // Grid::increase_hospital_size (grid.rs:233-238)
    g.hospital_expanded = Rect{g.hospital_initial.sx, g.hospital_initial.sy, G, G};
    const int max_x = std::max(std::max(g.hospital_initial.ex, g.hospital_expanded.ex), G);
    g.pitch = (uint32_t)max_x + 2u;      // one always-vacant padding column right of the last cell
    g.pitch = (g.pitch + 15u) & ~15u;  // rows start on 16-byte boundaries
    g.rows = (uint32_t)G + 1u;
    return g;
}

std::string validate_config(const epi_config& c) {
    auto pct = [](double p) { return p >= 0.0 && p <= 1.0; };
    if (c.number_of_agents == 0) return c.population_csv_file[0] ? "the population file holds no records (reference panics \"No citizens!\")"
                                                                 : "population.Auto.number_of_agents must be > 0 (reference panics \"No citizens!\")";
    if (c.grid_size < 10) return "geography_parameters.grid_size must be >= 10 (no offices otherwise)";
    if (c.grid_size > MAX_COORD - 1) return "geography_parameters.grid_size must be <= 16382 (cell packing)";
    if (!pct(c.public_transport_percentage) || !pct(c.working_percentage) || !pct(c.hospital_beds_percentage)) return "percentage out of [0,1]";
    if (!pct(c.regular_transmission_rate) || !pct(c.high_transmission_rate) || !pct(c.death_rate) ||
        !pct(c.percentage_asymptomatic_population) || !pct(c.percentage_severe_infected_population))
        return "disease percentage out of [0,1]";
    if (c.has_lockdown && !pct(c.essential_workers_population)) return "essential_workers_population out of [0,1]";
    if (c.n_vaccinations < 0 || c.n_vaccinations > EPI_MAX_VACCINATIONS) return "too many Vaccinate interventions";
    for (int i = 0; i < c.n_vaccinations; ++i)
        if (!pct(c.vaccinate_percent[i])) return "Vaccinate.percent out of [0,1]";
    const uint64_t infections = (uint64_t)c.exposed + c.infected_mild_asymptomatic + c.infected_mild_symptomatic + c.infected_severe;
    if (infections > c.number_of_agents) return "more starting infections than agents (citizen_factory.rs:113-115)";
    if (c.hours / 24u >= ST_DAY_MAX) return "hours too large for the 14-bit infection_day field";
    if (c.number_of_agents > (1u << 27)) return "number_of_agents must be <= 2^27";
    return "";
}

void apply_commute_plan(HostAgents& a, uint32_t n_agents, int region, const std::vector<uint32_t>& commute_row) {
    uint32_t i = 0;
    for (size_t to = 0; to < commute_row.size(); ++to) {
        uint32_t want = commute_row[to];
        while (want > 0 && i < n_agents) {
            const uint32_t s = a.st[i];
            const bool working = ((s >> ST_WS_SHIFT) & 3u) != WS_NA;
            if (working && ((a.reg[i] >> 8) & 0xFFu) == (uint32_t)region && (s & ST_PT)) {
                a.reg[i] = (a.reg[i] & 0xFFu) | ((uint32_t)to << 8);
                --want;
            }
            ++i;
        }
    }
}

Params make_params(const epi_config& c, const Geometry& g, uint64_t seed, int region) {
    Params P{};
    P.n = c.number_of_agents;
    P.grid_size = g.grid_size;
    P.pitch = g.pitch;
    P.rows = g.rows;
    P.zone[0] = g.transport; P.zone[1] = g.housing; P.zone[2] = g.hospital_resized; P.zone[3] = g.hospital_expanded;
    P.work = g.work;
    P.hospital_gen = 0;
    P.house_nx = g.house_nx; P.office_nx = g.office_nx;
    P.house_ny = g.house_ny; P.office_ny = g.office_ny;
    P.regular_start = c.regular_transmission_start_day;
    P.high_start = c.high_transmission_start_day;
    P.last_day = c.last_day;
    P.exposed_duration = c.exposed_duration;
    P.pre_symptomatic_duration = c.pre_symptomatic_duration;
    P.thr_rate[0] = 0;
    P.thr_rate[1] = bernoulli_threshold(c.regular_transmission_rate);
    P.thr_rate[2] = bernoulli_threshold(c.high_transmission_rate);
    P.thr_death = bernoulli_threshold(c.death_rate);
    P.thr_symptomatic = bernoulli_threshold(1.0 - c.percentage_asymptomatic_population);
    P.thr_severe = bernoulli_threshold(c.percentage_severe_infected_population);
    // Disease::is_to_be_hospitalized: current rate >= high rate (disease/mod.rs:97-99)
    P.hospitalize_mask = (0.0 >= c.high_transmission_rate ? 1u : 0u) | (c.regular_transmission_rate >= c.high_transmission_rate ? 2u : 0u) | 4u;
    uint32_t bits = 1;
    while ((1ull << bits) < (uint64_t)P.n) ++bits;
    P.id_bits = bits;
    P.seed = seed;
    for (uint32_t r = 0; r < 10; ++r) {
        P.rk[r][0] = (uint32_t)seed + r * 0x9E3779B9u;
        P.rk[r][1] = (uint32_t)(seed >> 32) + r * 0xBB67AE85u;
    }
    P.region = region;
    return P;
}

// ---- Population::Csv ---------------------------------------------------------------------------------------------------
namespace {
// one CSV record -> fields (RFC 4180 quoting, as the csv crate reads it by default)
bool split_csv_line(const std::string& line, std::vector<std::string>& out) {
    out.clear();
    std::string cur;
    bool quoted = false;
    for (size_t i = 0; i < line.size(); ++i) {
        const char ch = line[i];
        if (quoted) {
            if (ch == '"') {
                if (i + 1 < line.size() && line[i + 1] == '"') { cur.push_back('"'); ++i; }
                else quoted = false;
            } else cur.push_back(ch);
        } else if (ch == '"' && cur.empty()) quoted = true;
        else if (ch == ',') { out.push_back(cur); cur.clear(); }
        else cur.push_back(ch);
    }
    out.push_back(cur);
    return !quoted;
}
}  // namespace

PopulationRecords read_population_csv(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Could not read population file: " + path);  // grid.rs:202
    PopulationRecords r;
    std::string line;
    std::vector<std::string> fields;
    int col_ind = -1, col_age = -1, col_working = -1, col_pt = -1;
    size_t n_cols = 0, line_no = 0;
    auto fail = [&](const std::string& what) { throw std::runtime_error("Could not deserialize population line " + std::to_string(line_no) + ": " + what); };
    auto parse_bool = [&](const std::string& v) -> uint8_t {  // population_record.rs:34-43
        if (v == "True") return 1;
        if (v == "False") return 0;
        fail("invalid value: string \"" + v + "\", expected True or False");
        return 0;
    };
    while (std::getline(f, line)) {
        ++line_no;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;  // the csv crate skips empty lines
        if (!split_csv_line(line, fields)) fail("unterminated quoted field");
        if (n_cols == 0) {  // header row: serde maps the struct fields by column name, other columns are ignored
            n_cols = fields.size();
            for (size_t k = 0; k < fields.size(); ++k) {
                if (fields[k] == "ind") col_ind = (int)k;
                else if (fields[k] == "age") col_age = (int)k;
                else if (fields[k] == "working") col_working = (int)k;
                else if (fields[k] == "pub_transport") col_pt = (int)k;
            }
            if (col_ind < 0) fail("missing field `ind`");
            if (col_age < 0) fail("missing field `age`");
            if (col_working < 0) fail("missing field `working`");
            if (col_pt < 0) fail("missing field `pub_transport`");
            continue;
        }
        if (fields.size() != n_cols) fail("found record with " + std::to_string(fields.size()) + " fields, but the previous record has " + std::to_string(n_cols) + " fields");
        const std::string& ind = fields[(size_t)col_ind];  // ind: u32
        if (ind.empty() || ind.size() > 10 || ind.find_first_not_of("0123456789") != std::string::npos || std::stoull(ind) > 0xFFFFFFFFull) fail("invalid digit found in string (field `ind`)");
        r.working.push_back(parse_bool(fields[(size_t)col_working]));
        r.pub_transport.push_back(parse_bool(fields[(size_t)col_pt]));
    }
    return r;
}

epi_config resolve_population(const epi_config& cfg, PopulationRecords& records) {
    epi_config c = cfg;
    c.population_csv_file[EPI_PATH_MAX - 1] = 0;
    if (c.population_csv_file[0]) {
        records = read_population_csv(c.population_csv_file);
        if (records.size() > (size_t)(1u << 27)) throw std::runtime_error("population file: more than 2^27 records");
        c.number_of_agents = (uint32_t)records.size();
    }
    return c;
}

void build_population(const epi_config& c, const Geometry& g, uint64_t seed, int region, HostAgents& out, const PopulationRecords* records) {
    const uint32_t n = c.number_of_agents;
    if (records && records->size() != n) throw std::runtime_error("population records do not match number_of_agents");
    if (g.n_houses == 0 || g.n_offices == 0) throw std::runtime_error("grid too small: no houses or offices");
    if ((uint64_t)n > 4ull * g.n_houses)
        throw std::runtime_error(records ? "Cannot accommodate citizens into homes! There are " + std::to_string(n) + " citizens, but " + std::to_string(4ull * g.n_houses) + " home points"  // grid.rs:216-223
                                         : std::string("more than 4 agents per house: population does not fit the housing area (grid.rs:140-142)"));
    const uint32_t H = g.n_houses;
    const uint64_t thr_working = bernoulli_threshold(c.working_percentage);
    const uint64_t thr_pt = bernoulli_threshold(c.public_transport_percentage);
    const uint64_t thr_staff = bernoulli_threshold(kHospitalStaffPercentage);
    const uint64_t thr_essential = bernoulli_threshold(c.has_lockdown ? c.essential_workers_population : 0.0);
    // init draws are keyed on the CREATION number, so an agent's attributes do not depend on how agents are numbered
    auto draw = [&](uint32_t creation, uint32_t slot) { return philox_draw(seed, creation, 0, DOM_INIT, slot); };

    // Public-transport users are capped by the number of transport points Area::random_points can hand out
    // (grid.rs:96-101 "fix the hack", area.rs:64-74): ceil(sqrt(n as f32)) columns x rows clipped to the strip.
    const double want = (double)n * (c.public_transport_percentage + 0.1) * (c.working_percentage + 0.1);
    const size_t want_points = (size_t)std::ceil(want);
    const size_t side = (size_t)std::ceil(std::sqrt((float)want_points));
    const size_t tw = (size_t)(g.transport.ex - g.transport.sx + 1), th = (size_t)(g.transport.ey - g.transport.sy + 1);
    const size_t pt_capacity = std::min(want_points, std::min(side, tw) * std::min(side, th));

    // ---- the citizens in creation order (citizen_factory.rs:58-88 / Citizen::from_record citizen/mod.rs:155-180) ----
    HostAgents made;
    made.resize(n);
    size_t pt_users = 0;
    for (uint32_t cr = 0; cr < n; ++cr) {
        bool working, pt;
        if (records) {  // pub_transport is taken as it is
            working = records->working[cr] != 0;
            pt = records->pub_transport[cr] != 0;
        } else {
            working = bernoulli(draw(cr, IS_WORKING), thr_working);
            pt = bernoulli(draw(cr, IS_PT), thr_pt) && working && pt_users < pt_capacity;
        }
        if (pt) ++pt_users;
        uint32_t ws = WS_NA;
        if (working) {
            ws = bernoulli(draw(cr, IS_STAFF), thr_staff) ? WS_STAFF : WS_NORMAL;
            if (ws == WS_NORMAL && bernoulli(draw(cr, IS_ESSENTIAL), thr_essential)) ws = WS_ESSENTIAL;
        }
        const uint32_t immunity_plus2 = (uint32_t)mulhi64(draw(cr, IS_IMMUNITY), 5);
        made.st[cr] = ST_S | (immunity_plus2 << ST_IMM_SHIFT) | (pt ? ST_PT : 0u) | (ws << ST_WS_SHIFT) | (AK_HOME << ST_AREA_SHIFT);
        made.home[cr] = house_origin(g, cr % H);  // homes_iter.cycle() (grid.rs:108-113, 205-206)
        made.work[cr] = working ? office_origin(g, cr % g.n_offices) : 0u;
        made.wsa[cr] = ws == WS_STAFF ? kRoutineWorkTime : 0u;
        // start cell: the k housemates of a house take the first k of (sx,sy),(sx,sy+1),(sx+1,sy),(sx+1,sy+1) in creation
        // order; a lone occupant gets a uniformly random corner (area.rs:64-74)
        const uint32_t house = cr % H, rank = cr / H;
        const uint32_t housemates = (n - 1 - house) / H + 1;  // creation numbers house, house + H, ... < n
        const int sx = g.housing.sx + 2 * (int)(house % (uint32_t)g.house_nx);
        const int sy = g.housing.sy + 2 * (int)(house / (uint32_t)g.house_nx);
        int x, y;
        if (housemates == 1) {
            x = sx + (int)mulhi64(draw(cr, IS_STARTX), 2);
            y = sy + (int)mulhi64(draw(cr, IS_STARTY), 2);
        } else {
            x = sx + (int)(rank / 2);
            y = sy + (int)(rank % 2);
        }
        made.cell[cr] = ((uint32_t)y << CELL_BITS) | (uint32_t)x;
    }

    // ---- agent numbering: id -> creation number ----
    std::vector<uint32_t> creation_of(n);
    {
        const AgentOrder order = agent_order();
        uint32_t i = 0;
        if (order == AgentOrder::Creation) {
            for (; i < n; ++i) creation_of[i] = i;
        } else {
            for (uint32_t house = 0; house < H && house < n; ++house)
                for (uint32_t cr = house; cr < n; cr += H) creation_of[i++] = cr;
        }
        if (i != n) throw std::runtime_error("agent numbering: internal error");
    }
    out.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t cr = creation_of[i];
        out.cell[i] = made.cell[cr]; out.st[i] = made.st[cr]; out.t0[i] = 0; out.home[i] = made.home[cr]; out.work[i] = made.work[cr];
        out.wsa[i] = made.wsa[cr];
        out.reg[i] = (uint32_t)region | ((uint32_t)region << 8);
    }

    // starting infections: uniform without replacement, then exposed / asymptomatic / mild / severe in that order
    // (citizen_factory.rs:112-134)
    const uint32_t total = c.exposed + c.infected_mild_asymptomatic + c.infected_mild_symptomatic + c.infected_severe;
    std::vector<uint32_t> chosen;
    chosen.reserve(total);
    std::unordered_set<uint32_t> seen;
    for (uint32_t k = 0; chosen.size() < total; ++k) {
        const uint32_t idx = (uint32_t)mulhi64(philox_draw(seed, k, 0, DOM_STARTINF, 0), n);
        if (seen.insert(idx).second) chosen.push_back(idx);
    }
    size_t q = 0;
    auto infect = [&](uint32_t count, uint32_t state, uint32_t sev, uint32_t day) {
        for (uint32_t j = 0; j < count; ++j) {
            uint32_t& s = out.st[chosen[q++]];
            s = (s & ~ST_STATE_MASK) | state | (sev << ST_SEV_SHIFT) | (day << ST_DAY_SHIFT);
        }
    };
    infect(c.exposed, ST_E, 0, 0);  // Exposed { at_hour: 0 }
    infect(c.infected_mild_asymptomatic, ST_I, SEV_ASYM, 1);
    infect(c.infected_mild_symptomatic, ST_I, SEV_MILD, 1);
    infect(c.infected_severe, ST_I, SEV_SEVERE, 1);
}

}  // namespace epi
