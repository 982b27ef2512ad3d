// =============================================================================
// oracle/epi_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement (C++17) of the reference's per-hour agent step, following
// the reference source function by function (each function cites file:line under
// /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, link or call this.  The
// product (epirust_b200/csrc) never includes or links anything from oracle/.
//
// PARITY STATUS: the reference (Rust) cannot be compiled in this image (no
// cargo/rustc) and it draws every random number from an unseeded
// rand::thread_rng(), so no reference test pins a stochastic outcome.
//   * deterministic pieces are pinned against every known-answer test the
//     reference's own unit tests hold (tests/test_oracle_kat.py)
//   * stochastic outcomes are "parity unpinned": rand 0.8.5 (unvendored, no
//     Cargo.lock) primitives are restated from their published algorithm
//     (gen_bool: u64 < p*2^64; gen_range/choose: uniform integer) and only
//     distributional equivalence is claimed.
//
// Three draw sources share one code path (struct Rng below):
//   KEYED    Philox4x32-10 keyed on (seed, agent, hour, slot) -- the same slot
//            convention the CUDA kernels use, so whole runs are bit-identical
//            oracle <-> GPU.  Phase B commits in ascending agent id.
//   TABLE    draws injected from a caller-supplied table (16 u64 per agent).
//   STREAM   one sequential mt19937_64 stream per worker thread, consumed like
//            the reference consumes thread_rng (early exit, no slot meaning),
//            phase B in hash-map iteration order.  This is the "reference-like"
//            mode used for the ensemble CI and for the timed CPU baseline.
// =============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// common/src/models/custom_types.rs:22-27
typedef uint32_t Hour;
typedef uint32_t Count;
typedef uint32_t Day;
typedef uint32_t Size;
typedef int32_t CoOrdinate;
typedef double Percentage;

// engine/src/models/constants.rs:22-50
namespace constants {
const Percentage HOUSE_AREA_RELATIVE_SIZE = 0.4;
const Percentage TRANSPORT_AREA_RELATIVE_SIZE = 0.2;
const Percentage WORK_AREA_RELATIVE_SIZE = 0.2;
const Percentage INITIAL_HOSPITAL_RELATIVE_SIZE = 0.1;
const Hour NUMBER_OF_HOURS = 24;
const Hour ROUTINE_START_TIME = 0;
const Hour SLEEP_START_TIME = 1;
const Hour SLEEP_END_TIME = 6;
const Hour ROUTINE_TRAVEL_START_TIME = 7;
const Hour ROUTINE_WORK_TIME = 8;
const Hour ROUTINE_TRAVEL_END_TIME = 17;
const Hour ROUTINE_WORK_END_TIME = 16;
const Hour ROUTINE_END_TIME = 23;
const Hour NON_WORKING_TRAVEL_END_TIME = 12;
const Hour HOURS_IN_A_DAY = 24;
const Day QUARANTINE_DAYS = 14;
const int IMMUNITY_RANGE[5] = {-2, -1, 0, 1, 2};
const int RANGE_FOR_EXPOSED[3] = {-1, 0, 1};
const Percentage HOSPITAL_STAFF_PERCENTAGE = 0.002;
const Size HOME_SIZE = 2;
const Size OFFICE_SIZE = 10;
const Day ASYMPTOMATIC_LAST_DAY = 9;
const Day MILD_INFECTED_LAST_DAY = 12;
}  // namespace constants

// -----------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123 constants).  Independent copy:
// the product has its own in epirust_b200/csrc/philox.h.
// -----------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Draw domains (4th counter word) and per-agent-hour draw slots.  This is the
// (seed, agent, hour, draw) convention of BASELINE.json north_star; DESIGN.md
// section "Draw slots" is the normative statement, restated here and in the kernels.
//
// Hour-step draws (DOM_STEP) are laid out so that the common agent-hour needs ONE
// Philox4x32 block (counter = (agent, hour, block, DOM_STEP)):
//   block 0: word0 = PICK (u32)  word1 = FACTOR (u32)  words2,3 = A (u64)
//   block 1: word0 = PX (u32)    word1 = PY (u32)
//   block 2+(j>>1): EXPOSE j (u64) = words 0,1 for even j, words 2,3 for odd j   (j = 0..7)
// u32 draws feed uniform-integer choices (index = mulhi32(draw, n), like rand 0.8's
// 32-bit widening-multiply sampler without the negligible rejection zone), u64 draws
// feed Bernoulli trials (draw < p * 2^64, exactly rand 0.8's Bernoulli).
// Every other domain (init, vaccinate, ...) uses generic u64 slots: slot s = block s>>1,
// words 0,1 for even s, words 2,3 for odd s.
enum Domain : uint32_t { DOM_STEP = 0, DOM_INIT = 1, DOM_VACCINATE = 2, DOM_MIGRATE = 3, DOM_STARTINF = 4, DOM_ARRIVAL = 5 };
enum Slot : uint32_t {
    SLOT_PICK = 0,    // move_agent_from neighbour choose (citizen/mod.rs:430)                       u32
    SLOT_FACTOR = 1,  // on_exposed RANGE_FOR_EXPOSED.choose (default_disease_handler.rs:53)          u32
    SLOT_A = 2,       // on_exposed symptomatic / on_infected severe / is_to_be_deceased              u64
    SLOT_PX = 3,      // Area::get_random_point x   (area.rs:77)                                      u32
    SLOT_PY = 4,      // Area::get_random_point y   (area.rs:78)                                      u32
    SLOT_EXPOSE0 = 8  // +j : gen_bool(rate of Moore neighbour j), j=0..7 (default_disease_handler.rs:79)  u64
};
const int SLOTS_PER_AGENT = 16;
// init slots (DOM_INIT, hour word = 0)
enum InitSlot : uint32_t { IS_WORKING = 0, IS_PT = 1, IS_STAFF = 2, IS_IMMUNITY = 3, IS_ESSENTIAL = 4, IS_STARTX = 5, IS_STARTY = 6 };

inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

// rand 0.8 Bernoulli::new(p): p==1.0 -> always true, else p_int = (p * 2^64) as u64; sample: u64 < p_int
inline uint64_t bernoulli_threshold(double p) {
    if (p >= 1.0) return UINT64_MAX;  // sentinel "always"
    if (p <= 0.0) return 0;
    return (uint64_t)(p * 18446744073709551616.0);
}
inline bool bernoulli(uint64_t draw, uint64_t thr) { return draw < thr || thr == UINT64_MAX; }

struct Rng {
    enum Mode { KEYED, TABLE, STREAM } mode = KEYED;
    // KEYED
    uint64_t seed = 0;
    uint32_t agent = 0, hour = 0, domain = DOM_STEP;
    // TABLE
    const uint64_t* row = nullptr;
    // STREAM
    std::mt19937_64* stream = nullptr;

    void block(uint32_t b, uint32_t o[4]) const {
        uint32_t ctr[4] = {agent, hour, b, domain};
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32_10(ctr, key, o);
    }
    // a u64 draw
    uint64_t next(uint32_t slot) {
        switch (mode) {
            case KEYED: {
                uint32_t o[4];
                if (domain == DOM_STEP) {
                    if (slot == SLOT_A) { block(0, o); return (uint64_t)o[2] | ((uint64_t)o[3] << 32); }
                    if (slot < SLOT_EXPOSE0 || slot >= SLOT_EXPOSE0 + 8) throw std::logic_error("not a u64 step slot");
                    const uint32_t j = slot - SLOT_EXPOSE0;
                    block(2 + (j >> 1), o);
                    return (j & 1) ? ((uint64_t)o[2] | ((uint64_t)o[3] << 32)) : ((uint64_t)o[0] | ((uint64_t)o[1] << 32));
                }
                block(slot >> 1, o);
                return (slot & 1) ? ((uint64_t)o[2] | ((uint64_t)o[3] << 32)) : ((uint64_t)o[0] | ((uint64_t)o[1] << 32));
            }
            case TABLE: return row[slot];
            default: return (*stream)();
        }
    }
    // a u32 draw (hour-step domain only)
    uint32_t next32(uint32_t slot) {
        switch (mode) {
            case KEYED: {
                uint32_t o[4];
                switch (slot) {
                    case SLOT_PICK: block(0, o); return o[0];
                    case SLOT_FACTOR: block(0, o); return o[1];
                    case SLOT_PX: block(1, o); return o[0];
                    case SLOT_PY: block(1, o); return o[1];
                    default: throw std::logic_error("not a u32 step slot");
                }
            }
            case TABLE: return (uint32_t)row[slot];
            default: return (uint32_t)((*stream)() >> 32);
        }
    }
    // rand::Rng::gen_bool
    bool gen_bool(uint32_t slot, double p) {
        uint64_t thr = bernoulli_threshold(p);
        if (thr == UINT64_MAX) return true;  // Bernoulli ALWAYS_TRUE consumes no draw
        return next(slot) < thr;
    }
    // rand::Rng::gen_range(a..=b) : uniform integer.  64-bit flavour (init domains) and 32-bit flavour (hour step)
    int gen_range_incl(uint32_t slot, int a, int b) { return a + (int)mulhi64(next(slot), (uint64_t)(b - a + 1)); }
    int gen_range_incl32(uint32_t slot, int a, int b) { return a + (int)mulhi32(next32(slot), (uint32_t)(b - a + 1)); }
    // SliceRandom::choose / IteratorRandom::choose over n candidates : uniform index
    uint32_t choose_index(uint32_t slot, uint32_t n) { return (uint32_t)mulhi64(next(slot), n); }
    uint32_t choose_index32(uint32_t slot, uint32_t n) { return mulhi32(next32(slot), n); }
};

// -----------------------------------------------------------------------------
// geography/point.rs
// -----------------------------------------------------------------------------
struct Point {
    CoOrdinate x, y;
    bool operator==(const Point& o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point& o) const { return !(*this == o); }
};
// point.rs:59 -- neighbour offsets in iterator order
static const int NEIGHBOR_OFFSETS[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};

// -----------------------------------------------------------------------------
// geography/area.rs
// -----------------------------------------------------------------------------
struct Area {
    int location_id = 0;  // region index (reference: s16 region name)
    Point start_offset{0, 0}, end_offset{0, 0};
    // area.rs:37-41 equality ignores location_id
    bool operator==(const Area& o) const { return start_offset == o.start_offset && end_offset == o.end_offset; }
    bool operator!=(const Area& o) const { return !(*this == o); }
    // area.rs:83-88 inclusive on both ends
    bool contains(const Point& p) const {
        return start_offset.x <= p.x && end_offset.x >= p.x && start_offset.y <= p.y && end_offset.y >= p.y;
    }
    // area.rs:56-58 : Moore neighbours (fixed order) that lie inside the area
    int get_neighbors_of(Point p, Point out[8]) const {
        int n = 0;
        for (int j = 0; j < 8; ++j) {
            Point q{p.x + NEIGHBOR_OFFSETS[j][0], p.y + NEIGHBOR_OFFSETS[j][1]};
            if (contains(q)) out[n++] = q;
        }
        return n;
    }
    // area.rs:76-81
    Point get_random_point(Rng& rng) const {
        int rx = rng.gen_range_incl32(SLOT_PX, start_offset.x, end_offset.x);
        int ry = rng.gen_range_incl32(SLOT_PY, start_offset.y, end_offset.y);
        return Point{rx, ry};
    }
    // area.rs:90-92 (sic: (ex-sx)*(ey-sy), not the inclusive cell count)
    Count get_number_of_cells() const {
        return (Count)((end_offset.x - start_offset.x) * (end_offset.y - start_offset.y));
    }
    // area.rs:119-146 AreaIterator: x fastest, then y, inclusive
    template <class F> bool iter_find(F pred, Point& found) const {
        for (int y = start_offset.y; y <= end_offset.y; ++y)
            for (int x = start_offset.x; x <= end_offset.x; ++x)
                if (pred(Point{x, y})) { found = Point{x, y}; return true; }
        return false;
    }
    std::vector<Point> iter_all() const {
        std::vector<Point> v;
        for (int y = start_offset.y; y <= end_offset.y; ++y)
            for (int x = start_offset.x; x <= end_offset.x; ++x) v.push_back(Point{x, y});
        return v;
    }
};

// area.rs:95-117
inline std::vector<Area> area_factory(Point start_point, Point end_point, uint32_t size, int region) {
    int fx = (end_point.x - start_point.x + 1) / (int)size;
    int fy = (end_point.y - start_point.y + 1) / (int)size;
    std::vector<Area> areas;
    areas.reserve((size_t)std::max(0, fx) * (size_t)std::max(0, fy));
    Point cur = start_point;
    for (int i = 0; i < fy; ++i) {
        for (int j = 0; j < fx; ++j) {
            Area a;
            a.location_id = region;
            a.start_offset = cur;
            a.end_offset = Point{cur.x + (int)size - 1, cur.y + (int)size - 1};
            areas.push_back(a);
            cur.x += (int)size;
        }
        cur.x = start_point.x;
        cur.y += (int)size;
    }
    return areas;
}

// -----------------------------------------------------------------------------
// common/src/disease/mod.rs
// -----------------------------------------------------------------------------
struct Disease {
    Day regular_transmission_start_day = 0, high_transmission_start_day = 0, last_day = 0;
    Day asymptomatic_last_day = 0, mild_infected_last_day = 0;  // parsed, never used (constants.rs:48-50)
    Percentage regular_transmission_rate = 0, high_transmission_rate = 0, death_rate = 0;
    Percentage percentage_asymptomatic_population = 0, percentage_severe_infected_population = 0;
    Hour exposed_duration = 0, pre_symptomatic_duration = 0;
    // disease/mod.rs:88-95
    Percentage get_current_transmission_rate(Day infection_day) const {
        if (regular_transmission_start_day < infection_day && infection_day <= high_transmission_start_day)
            return regular_transmission_rate;
        else if (high_transmission_start_day < infection_day && infection_day <= last_day)
            return high_transmission_rate;
        return 0.0;
    }
    // disease/mod.rs:97-99
    bool is_to_be_hospitalized(Day infection_day) const {
        return get_current_transmission_rate(infection_day) >= high_transmission_rate;
    }
    // disease/mod.rs:105-107
    bool is_to_be_deceased(Rng& rng) const { return rng.gen_bool(SLOT_A, death_rate); }
};

// -----------------------------------------------------------------------------
// state_machine/state.rs
// -----------------------------------------------------------------------------
enum StateKind : uint8_t { Susceptible = 0, Exposed = 1, Infected = 2, Recovered = 3, Deceased = 4 };
enum SeverityKind : uint8_t { Pre = 0, Asymptomatic = 1, Mild = 2, Severe = 3 };
struct State {
    StateKind kind = Susceptible;
    SeverityKind severity = Pre;
    Hour at_hour = 0;        // Exposed{at_hour} or Infected{Pre{at_hour}}
    Day infection_day = 0;   // Infected only
    bool operator==(const State& o) const {
        if (kind != o.kind) return false;
        if (kind == Exposed) return at_hour == o.at_hour;
        if (kind == Infected)
            return infection_day == o.infection_day && severity == o.severity && (severity != Pre || at_hour == o.at_hour);
        return true;
    }
    bool is_mild_symptomatic() const { return kind == Infected && severity == Mild; }   // state.rs:45
    bool is_infected_severe() const { return kind == Infected && severity == Severe; }  // state.rs:49
    void update_infection_day() { if (kind == Infected) infection_day += 1; }           // state.rs:69-73
    static State infected(Day d, SeverityKind s, Hour at = 0) { State x; x.kind = Infected; x.severity = s; x.infection_day = d; x.at_hour = at; return x; }
    static State expose(Hour at) { State x; x.kind = Exposed; x.at_hour = at; return x; }
    static State simple(StateKind k) { State x; x.kind = k; return x; }
};

enum WorkStatusKind : uint8_t { Normal = 0, Essential = 1, HospitalStaff = 2, NA = 3 };  // citizen/work_status.rs:22-28

struct Citizen;
struct CitizenLocationMap;

// -----------------------------------------------------------------------------
// disease_state_machine.rs (struct) -- handlers are in Disease-based free functions below
// -----------------------------------------------------------------------------
struct DiseaseStateMachine {
    State state;
    Day get_infection_day() const { return state.kind == Infected ? state.infection_day : 0; }  // :37-42
    bool is_susceptible() const { return state.kind == Susceptible; }
    bool is_infected() const { return state.kind == Infected; }
    bool is_symptomatic() const { return state.is_mild_symptomatic() || state.is_infected_severe(); }  // :100-102
    bool is_deceased() const { return state.kind == Deceased; }
    void increment_infection_day() { state.update_infection_day(); }
};

// -----------------------------------------------------------------------------
// geography/grid.rs + geography/mod.rs
// -----------------------------------------------------------------------------
struct Grid {
    Size grid_size = 0;
    Area housing_area, work_area, transport_area, hospital_area;
    std::vector<Area> houses, offices;
    // houses_occupancy / offices_occupancy heaps (grid.rs:279-341) live in travel code (oracle/epi_oracle_travel.hpp)
    std::vector<uint32_t> house_occupants, office_occupants;

    // grid.rs:233-238
    void increase_hospital_size(Size gs) {
        hospital_area.end_offset = Point{(CoOrdinate)gs, (CoOrdinate)gs};
    }
    // grid.rs:240-261
    void resize_hospital(int number_of_agents, double hospital_staff_percentage, double hospital_beds_percentage) {
        Count hospital_bed_count =
            (Count)std::ceil((double)number_of_agents * hospital_beds_percentage + (double)number_of_agents * hospital_staff_percentage);
        if (hospital_bed_count <= hospital_area.get_number_of_cells()) {
            CoOrdinate hospital_end_y =
                (CoOrdinate)(hospital_bed_count / (uint32_t)(hospital_area.end_offset.x - hospital_area.start_offset.x));
            hospital_area.end_offset = Point{hospital_area.end_offset.x, hospital_end_y};
        }
    }
};

// geography/mod.rs:33-70
inline Grid define_geography(Size grid_size, int region) {
    int home_width = (int)std::ceil((double)grid_size * constants::HOUSE_AREA_RELATIVE_SIZE);
    int transport_start = home_width;
    int transport_end = home_width + (int)std::ceil((double)grid_size * constants::TRANSPORT_AREA_RELATIVE_SIZE);
    int work_area_start = transport_end;
    int work_area_end = transport_end + (int)std::ceil((double)grid_size * constants::WORK_AREA_RELATIVE_SIZE);
    int hospital_start = work_area_end;
    int hospital_end = work_area_end + (int)std::ceil((double)grid_size * constants::INITIAL_HOSPITAL_RELATIVE_SIZE);
    Grid g;
    g.grid_size = grid_size;
    auto mk = [&](int sx, int ex) { Area a; a.location_id = region; a.start_offset = Point{sx, 0}; a.end_offset = Point{ex, (CoOrdinate)grid_size}; return a; };
    g.housing_area = mk(0, home_width - 1);
    g.transport_area = mk(transport_start, transport_end - 1);
    g.work_area = mk(work_area_start, work_area_end - 1);
    g.hospital_area = mk(hospital_start, hospital_end - 1);
    g.houses = area_factory(g.housing_area.start_offset, g.housing_area.end_offset, constants::HOME_SIZE, region);
    g.offices = area_factory(g.work_area.start_offset, g.work_area.end_offset, constants::OFFICE_SIZE, region);
    return g;
}

// -----------------------------------------------------------------------------
// citizen/mod.rs
// -----------------------------------------------------------------------------
struct Citizen {
    uint32_t id = 0;  // reference: Uuid v4; here the creation index (the phase-B priority)
    int immunity = 0;
    Area home_location, work_location;
    bool vaccinated = false, uses_public_transport = false, hospitalized = false;
    Point transport_location{0, 0};
    DiseaseStateMachine state_machine;
    bool isolated = false;
    Area current_area;
    WorkStatusKind work_status = NA;
    Hour work_start_at = 0;  // payload of WorkStatus::HospitalStaff
    bool work_quarantined = false;

    bool is_hospital_staff() const { return work_status == HospitalStaff; }  // :468
    bool is_working() const { return work_status != NA; }                    // :480
    bool is_essential_worker() const { return work_status == Essential; }    // :484
    // :452-454
    bool can_move() const {
        return !(state_machine.is_symptomatic() || hospitalized || state_machine.is_deceased() || isolated);
    }
    // :182-185 (day + immunity) as u32, wrapping
    Percentage get_infection_transmission_rate(const Disease& d) const {
        return d.get_current_transmission_rate((Day)((int32_t)state_machine.get_infection_day() + immunity));
    }

    Point routine(Point cell, Hour simulation_hour, const Grid& grid, const CitizenLocationMap& map, Rng& rng, const Disease& dh);
    Point perform_movements(Point cell, Hour hour_of_day, Hour simulation_hr, const Grid& grid, const CitizenLocationMap& map, Rng& rng,
                            const Disease& dh);
    Point hospitalize(Point cell, const Area& hospital, const CitizenLocationMap& map, Rng& rng, const Disease& dh);
    Point goto_area(const Area& target_area, const CitizenLocationMap& map, Point cell, Rng& rng) const;
    Point deceased(const CitizenLocationMap& map, Point cell, Rng& rng, const Disease& dh);
    Point move_agent_from(const CitizenLocationMap& map, Point cell, Rng& rng) const;
    void update_infection_dynamics(Point cell, const CitizenLocationMap& map, Hour sim_hr, Rng& rng, const Disease& dh);
};

// -----------------------------------------------------------------------------
// models/events/counts.rs
// -----------------------------------------------------------------------------
struct Counts {
    Hour hour = 0;
    Count susceptible = 0, exposed = 0, infected = 0, hospitalized = 0, recovered = 0, deceased = 0;
    // counts.rs:126-140
    void update_counts(const Citizen& c) {
        switch (c.state_machine.state.kind) {
            case Susceptible: susceptible += 1; break;
            case Exposed: exposed += 1; break;
            case Infected: if (c.hospitalized) hospitalized += 1; else infected += 1; break;
            case Recovered: recovered += 1; break;
            case Deceased: deceased += 1; break;
        }
    }
    void clear() { susceptible = exposed = infected = hospitalized = recovered = deceased = 0; }  // :142-149
    Count total() const { return susceptible + exposed + infected + hospitalized + recovered + deceased; }  // :151-153
    void increment_hour() { hour += 1; }
};

// -----------------------------------------------------------------------------
// Point -> Citizen map (reference: FnvHashMap<Point, Citizen>, allocation_map.rs:44-49).
// Open addressing, FNV-1a over the 8 key bytes like the fnv crate.
// -----------------------------------------------------------------------------
class PointMap {
  public:
    std::vector<uint8_t> used;
    std::vector<Point> keys;
    std::vector<Citizen> vals;
    size_t mask = 0, count = 0;
    void init(size_t expected) {
        size_t cap = 64;
        while (cap < expected * 2 + 16) cap <<= 1;
        used.assign(cap, 0); keys.resize(cap); vals.resize(cap);
        mask = cap - 1; count = 0;
    }
    static uint64_t hash(const Point& p) {
        uint64_t h = 0xcbf29ce484222325ull;
        uint32_t w[2] = {(uint32_t)p.x, (uint32_t)p.y};
        const uint8_t* b = (const uint8_t*)w;
        for (int i = 0; i < 8; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
        return h ^ (h >> 29);  // fold high bits down before masking
    }
    size_t capacity() const { return used.size(); }
    void clear() { std::fill(used.begin(), used.end(), 0); count = 0; }
    void grow_if_needed() {
        if ((count + 1) * 10 < capacity() * 7) return;
        PointMap n; n.init(count * 2 + 16);
        for (size_t i = 0; i < capacity(); ++i) if (used[i]) n.insert(keys[i], vals[i]);
        *this = std::move(n);
    }
    long find_slot(const Point& p) const {
        size_t i = hash(p) & mask;
        while (used[i]) { if (keys[i] == p) return (long)i; i = (i + 1) & mask; }
        return -1;
    }
    bool contains_key(const Point& p) const { return find_slot(p) >= 0; }
    const Citizen* get(const Point& p) const { long s = find_slot(p); return s < 0 ? nullptr : &vals[s]; }
    Citizen* get_mut(const Point& p) { long s = find_slot(p); return s < 0 ? nullptr : &vals[s]; }
    // HashMap::entry(k).or_insert(v): returns the value now stored under k
    Citizen& entry_or_insert(const Point& p, const Citizen& c) {
        grow_if_needed();
        size_t i = hash(p) & mask;
        while (used[i]) { if (keys[i] == p) return vals[i]; i = (i + 1) & mask; }
        used[i] = 1; keys[i] = p; vals[i] = c; ++count;
        return vals[i];
    }
    // HashMap::insert: overwrites; returns true if a previous value existed
    bool insert(const Point& p, const Citizen& c) {
        grow_if_needed();
        size_t i = hash(p) & mask;
        while (used[i]) { if (keys[i] == p) { vals[i] = c; return true; } i = (i + 1) & mask; }
        used[i] = 1; keys[i] = p; vals[i] = c; ++count;
        return false;
    }
    bool remove(const Point& p, Citizen* out = nullptr) {
        long s = find_slot(p);
        if (s < 0) return false;
        if (out) *out = vals[s];
        size_t i = (size_t)s;
        // backward-shift deletion
        size_t j = i;
        for (;;) {
            j = (j + 1) & mask;
            if (!used[j]) break;
            size_t k = hash(keys[j]) & mask;
            bool in_between = (i <= j) ? (i < k && k <= j) : (i < k || k <= j);
            if (in_between) continue;
            keys[i] = keys[j]; vals[i] = vals[j]; i = j;
        }
        used[i] = 0; --count;
        return true;
    }
    size_t len() const { return count; }
};

struct Update {  // one element of `updates` in allocation_map.rs:82-92
    Point old_cell, new_cell;
    Citizen agent;
    bool infection_status;
};

// -----------------------------------------------------------------------------
// allocation_map.rs
// -----------------------------------------------------------------------------
struct CitizenLocationMap {
    Grid grid;
    PointMap current_locations, upcoming_locations;
    // cache of goto_hospital's hospital_area.iter().find(vacant) for the current hour: the
    // start-of-hour map is immutable during phase A so every caller would find the same cell.
    mutable bool hospital_cache_valid = false, hospital_cache_found = false;
    mutable Point hospital_cache_cell{0, 0};

    void init(const Grid& g, const std::vector<Citizen>& agents, const std::vector<Point>& points) {  // :52-65
        grid = g;
        current_locations.init(agents.size());
        upcoming_locations.init(agents.size());
        for (size_t i = 0; i < agents.size(); ++i) current_locations.insert(points[i], agents[i]);
    }
    // :136-142
    Point move_agent(Point old_cell, Point new_cell) const { return is_cell_vacant(new_cell) ? new_cell : old_cell; }
    // :152-154
    const Citizen* get_agent_for(const Point& cell) const { return current_locations.get(cell); }
    // :156-159
    bool is_point_in_grid(const Point& p) const {
        CoOrdinate e = (CoOrdinate)grid.grid_size;
        return p.x >= 0 && p.y >= 0 && p.x < e && p.y < e;
    }
    // :161-163
    bool is_cell_vacant(const Point& cell) const { return !current_locations.contains_key(cell); }
    Count current_population() const { return (Count)current_locations.len(); }
    // :144-150.  Returns (is_hospitalized, new_location)
    std::pair<bool, Point> goto_hospital(const Area& hospital_area, Point cell, Citizen& citizen, Rng& rng) const {
        bool found; Point x{0, 0};
        if (hospital_cache_valid) { found = hospital_cache_found; x = hospital_cache_cell; }
        else found = hospital_area.iter_find([&](Point p) { return is_cell_vacant(p); }, x);
        if (found) return {true, move_agent(cell, x)};
        // reference draws from a fresh RandomWrapper here; any iid source is equivalent
        return {false, move_agent(cell, citizen.home_location.get_random_point(rng))};
    }
    void prime_hospital_cache() const {
        hospital_cache_found = grid.hospital_area.iter_find([&](Point p) { return is_cell_vacant(p); }, hospital_cache_cell);
        hospital_cache_valid = true;
    }
    // :131-134
    void swap() { current_locations.clear(); std::swap(current_locations, upcoming_locations); }

    // :349-356
    void lock_city() {
        for (size_t i = 0; i < current_locations.capacity(); ++i)
            if (current_locations.used[i] && !current_locations.vals[i].is_essential_worker()) current_locations.vals[i].isolated = true;
    }
    // :358-365
    void unlock_city() {
        for (size_t i = 0; i < current_locations.capacity(); ++i)
            if (current_locations.used[i] && current_locations.vals[i].isolated) current_locations.vals[i].isolated = false;
    }
};

// -----------------------------------------------------------------------------
// state_machine/default_disease_handler.rs (impl DiseaseHandler for Disease)
// -----------------------------------------------------------------------------
// :32-39
inline bool is_to_be_hospitalize(const Disease& d, const State& s, int immunity) {
    if (s.kind == Infected && s.severity == Severe) return d.is_to_be_hospitalized((Day)((int32_t)s.infection_day + immunity));
    return false;
}
// :41-50
inline bool on_infected(const Disease& d, Hour sim_hr, const State& cur, Rng& rng, State& out) {
    if (cur.severity == Pre && sim_hr - cur.at_hour >= d.pre_symptomatic_duration) {
        bool is_severe = rng.gen_bool(SLOT_A, d.percentage_severe_infected_population);
        out = State::infected(cur.infection_day, is_severe ? Severe : Mild);
        return true;
    }
    return false;
}
// :52-62
inline bool on_exposed(const Disease& d, Hour at_hour, Hour sim_hr, Rng& rng, State& out) {
    int random_factor = constants::RANGE_FOR_EXPOSED[rng.choose_index32(SLOT_FACTOR, 3)];
    if (sim_hr - at_hour >= (Hour)((int32_t)d.exposed_duration + random_factor)) {
        bool symptoms = rng.gen_bool(SLOT_A, 1.0 - d.percentage_asymptomatic_population);
        out = symptoms ? State::infected(0, Pre, sim_hr) : State::infected(0, Asymptomatic);
        return true;
    }
    return false;
}
// :64-86
inline bool on_susceptible(const Disease& d, Hour sim_hr, Point cell, const Citizen& citizen, const CitizenLocationMap& map, Rng& rng, State& out) {
    if (!citizen.work_quarantined && !citizen.vaccinated) {
        for (int j = 0; j < 8; ++j) {
            Point p{cell.x + NEIGHBOR_OFFSETS[j][0], cell.y + NEIGHBOR_OFFSETS[j][1]};
            if (!citizen.current_area.contains(p)) continue;   // get_neighbors_of
            if (!map.is_point_in_grid(p)) continue;
            const Citizen* n = map.get_agent_for(p);
            if (!n) continue;
            if (!(n->state_machine.is_infected() && !n->hospitalized)) continue;
            if (rng.gen_bool(SLOT_EXPOSE0 + j, n->get_infection_transmission_rate(d))) {  // .find(): first success ends the scan
                out = State::expose(sim_hr);
                return true;
            }
        }
    }
    return false;
}
// :88-103
inline bool on_routine_end(const Disease& d, const State& cur, Rng& rng, State& out) {
    if (cur.kind == Infected) {
        if (cur.severity == Asymptomatic && cur.infection_day == constants::ASYMPTOMATIC_LAST_DAY) { out = State::simple(Recovered); return true; }
        if (cur.severity == Mild && cur.infection_day == constants::MILD_INFECTED_LAST_DAY) { out = State::simple(Recovered); return true; }
        if (cur.severity == Severe && cur.infection_day == d.last_day) {
            out = d.is_to_be_deceased(rng) ? State::simple(Deceased) : State::simple(Recovered);
            return true;
        }
    }
    return false;
}
// disease_state_machine.rs:53-70
inline State dsm_next(const DiseaseStateMachine& m, Hour sim_hr, Point cell, const Citizen& citizen, const CitizenLocationMap& map, Rng& rng,
                      const Disease& d) {
    State out;
    switch (m.state.kind) {
        case Susceptible: return on_susceptible(d, sim_hr, cell, citizen, map, rng, out) ? out : m.state;
        case Exposed: return on_exposed(d, m.state.at_hour, sim_hr, rng, out) ? out : m.state;
        case Infected: return on_infected(d, sim_hr, m.state, rng, out) ? out : m.state;
        default: return m.state;
    }
}

// ---- impl Citizen (citizen/mod.rs) -------------------------------------------
// :194-203
inline void Citizen::update_infection_dynamics(Point cell, const CitizenLocationMap& map, Hour sim_hr, Rng& rng, const Disease& dh) {
    state_machine.state = dsm_next(state_machine, sim_hr, cell, *this, map, rng, dh);
}
// :227-255
inline Point Citizen::routine(Point cell, Hour simulation_hour, const Grid& grid, const CitizenLocationMap& map, Rng& rng, const Disease& dh) {
    Point new_cell = cell;
    Hour current_hour = simulation_hour % constants::NUMBER_OF_HOURS;
    if (current_hour == constants::ROUTINE_START_TIME) {
        state_machine.increment_infection_day();
        new_cell = hospitalize(cell, grid.hospital_area, map, rng, dh);
    } else if (current_hour >= constants::SLEEP_START_TIME && current_hour <= constants::SLEEP_END_TIME) {
        if (!is_hospital_staff()) current_area = home_location;
    } else if (current_hour == constants::ROUTINE_END_TIME) {
        new_cell = deceased(map, cell, rng, dh);
    } else {
        new_cell = perform_movements(cell, current_hour, simulation_hour, grid, map, rng, dh);
    }
    return new_cell;
}
// :257-349
inline Point Citizen::perform_movements(Point cell, Hour hour_of_day, Hour simulation_hr, const Grid& grid, const CitizenLocationMap& map,
                                        Rng& rng, const Disease& dh) {
    Point new_cell = cell;
    switch (work_status) {
        case Normal:
        case Essential: {
            if (hour_of_day == constants::ROUTINE_TRAVEL_START_TIME || hour_of_day == constants::ROUTINE_TRAVEL_END_TIME) {
                if (uses_public_transport) {
                    new_cell = goto_area(grid.transport_area, map, cell, rng);
                    current_area = grid.transport_area;
                } else {
                    new_cell = move_agent_from(map, cell, rng);
                }
            } else if (hour_of_day == constants::ROUTINE_WORK_TIME) {
                new_cell = goto_area(work_location, map, cell, rng);
                current_area = work_location;
            } else if (hour_of_day == constants::ROUTINE_WORK_END_TIME) {
                new_cell = goto_area(home_location, map, cell, rng);
                current_area = home_location;
            } else {
                new_cell = move_agent_from(map, cell, rng);
            }
            update_infection_dynamics(new_cell, map, simulation_hr, rng, dh);
            break;
        }
        case HospitalStaff: {
            Hour since = simulation_hr >= work_start_at ? simulation_hr - work_start_at : 0;  // saturating_sub
            if (since == constants::HOURS_IN_A_DAY * constants::QUARANTINE_DAYS) {
                work_quarantined = true;
                return new_cell;
            }
            if (since == constants::HOURS_IN_A_DAY * constants::QUARANTINE_DAYS * 2) {
                new_cell = goto_area(home_location, map, cell, rng);
                current_area = home_location;
                work_start_at = simulation_hr + constants::HOURS_IN_A_DAY * constants::QUARANTINE_DAYS;
                return new_cell;
            }
            if (hour_of_day == constants::ROUTINE_WORK_TIME) {
                if (current_area != grid.hospital_area && work_start_at <= simulation_hr) {
                    new_cell = goto_area(grid.hospital_area, map, cell, rng);
                    current_area = grid.hospital_area;
                    work_start_at = simulation_hr;
                }
                work_quarantined = false;
            } else if (hour_of_day == constants::ROUTINE_WORK_END_TIME) {
                work_quarantined = true;
            } else {
                if (!work_quarantined && can_move()) new_cell = move_agent_from(map, cell, rng);
            }
            update_infection_dynamics(new_cell, map, simulation_hr, rng, dh);
            break;
        }
        case NA: {
            if (hour_of_day == constants::ROUTINE_WORK_TIME) {
                new_cell = goto_area(grid.housing_area, map, cell, rng);
                current_area = grid.housing_area;
            } else if (hour_of_day == constants::NON_WORKING_TRAVEL_END_TIME) {
                new_cell = goto_area(home_location, map, cell, rng);
                current_area = home_location;
            } else {
                new_cell = move_agent_from(map, cell, rng);
            }
            update_infection_dynamics(new_cell, map, simulation_hr, rng, dh);
            break;
        }
    }
    return new_cell;
}
// :351-365
inline Point Citizen::hospitalize(Point cell, const Area& hospital, const CitizenLocationMap& map, Rng& rng, const Disease& dh) {
    Point new_cell = cell;
    if (!hospitalized && is_to_be_hospitalize(dh, state_machine.state, immunity)) {
        auto r = map.goto_hospital(hospital, cell, *this, rng);
        new_cell = r.second;
        hospitalized = r.first;
    }
    return new_cell;
}
// :367-395
inline Point Citizen::goto_area(const Area& target_area, const CitizenLocationMap& map, Point cell, Rng& rng) const {
    bool override_movement = false;
    if (work_status == Normal || work_status == Essential) {
        if (work_location.contains(cell) && target_area == home_location &&
            (state_machine.state.is_mild_symptomatic() || state_machine.state.is_infected_severe()))
            override_movement = true;
    }
    if (!can_move() && !override_movement) return cell;
    if (is_working()) {
        Point new_cell = target_area.get_random_point(rng);
        if (!map.is_cell_vacant(new_cell)) new_cell = cell;
        return map.move_agent(cell, new_cell);
    }
    return move_agent_from(map, cell, rng);
}
// :397-413 ; DiseaseStateMachine::decease disease_state_machine.rs:72-77
inline Point Citizen::deceased(const CitizenLocationMap& map, Point cell, Rng& rng, const Disease& dh) {
    Point new_cell = cell;
    State out;
    if (on_routine_end(dh, state_machine.state, rng, out)) state_machine.state = out;
    if (state_machine.state.kind == Recovered) new_cell = map.move_agent(cell, home_location.get_random_point(rng));
    if ((state_machine.state.kind == Recovered || state_machine.state.kind == Deceased) && hospitalized) hospitalized = false;
    return new_cell;
}
// :415-432
inline Point Citizen::move_agent_from(const CitizenLocationMap& map, Point cell, Rng& rng) const {
    if (!can_move()) return cell;
    Point current_location = cell;
    if (!current_area.contains(cell)) current_location = current_area.get_random_point(rng);
    Point nb[8], cand[8];
    int n = current_area.get_neighbors_of(current_location, nb);
    int k = 0;
    for (int j = 0; j < n; ++j)
        if (map.is_point_in_grid(nb[j]) && map.is_cell_vacant(nb[j])) cand[k++] = nb[j];
    Point new_cell = cell;  // .unwrap_or(cell)
    if (k > 0) new_cell = cand[rng.choose_index32(SLOT_PICK, (uint32_t)k)];
    return map.move_agent(cell, new_cell);
}

}  // namespace orc
