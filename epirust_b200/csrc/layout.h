// Device data layout of one region engine.  Everything here is resident in HBM for the whole run.
//
// Per agent (structure of arrays, index = agent id = creation order, engine/src/citizen/citizen_factory.rs:44-47):
//   cell[N] u32   packed (y << 14) | x of the cell the agent stands on          (reference: the map key Point)
//   st[N]   u32   packed Citizen fields, layout below                          (citizen/mod.rs:48-63)
//   t0[N]   u32   at_hour of State::Exposed / Severity::Pre                     (state_machine/state.rs:22-37)
//   home[N] u32   packed (sy << 14) | sx origin of the agent's 2x2 house          (Citizen.home_location)
//   work[N] u32   packed origin of the agent's 10x10 office (unused for WorkStatus::NA)  (Citizen.work_location)
//                 (origins instead of indices: no div/mod in the hot kernel; the C ABI speaks house/office indices)
//   wsa[N]  u32   WorkStatus::HospitalStaff.work_start_at (0.14 % of agents)    (citizen/work_status.rs:26)
//   prop[N] u32   this hour's proposal, hour kernel -> commit kernel
// Per cell, row-major (y-major, so x neighbours are adjacent bytes and the houses of consecutive agents are adjacent in
// memory): pitch >= grid_size + 2 bytes per row, rows = grid_size + 1 (Area ends are inclusive).  The allocation carries
// GRID_YOFF always-zero rows above and below and GRID_XOFF zero bytes in front, so the 5x5 window around any cell can be
// loaded without bounds checks; cell_offset() below is the only place that knows the layout.  (A column-strip-major
// layout -- 4-cell-wide strips, rows consecutive inside a strip, 2-4 sectors per window instead of 5-10 -- was measured
// 15 % slower overall: it halves the L2 sector requests of the random work-hour windows but breaks the coalescing of the
// home hours, where consecutive agents live in x-adjacent houses; DRAM bytes did not change.)
//   grid[cells]  u8   0 vacant, 1 occupied & not infectious, 2 occupied & regular-rate, 3 occupied & high-rate
//                     (replaces AgentLocationMap / FnvHashMap<Point, Citizen>, allocation_map.rs:44-49: vacancy and
//                      the neighbour's transmission rate are the only things other agents read from a cell)
//   claim[cells] u32  (stamp << id_bits) | (id_mask - agent id), atomicMax: lowest agent id wins the cell this hour
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace epi {

// ---- agent state word ------------------------------------------------------------------------------------------
enum : uint32_t {
    ST_S = 0, ST_E = 1, ST_I = 2, ST_R = 3, ST_D = 4,               // state_machine/state.rs:31-37
    ST_ABSENT = 7,                                                   // empty agent slot (multi-region engines: travellers leave and arrive)
    SEV_PRE = 0, SEV_ASYM = 1, SEV_MILD = 2, SEV_SEVERE = 3,         // state_machine/state.rs:22-28
    WS_NORMAL = 0, WS_ESSENTIAL = 1, WS_STAFF = 2, WS_NA = 3,        // citizen/work_status.rs:22-28
    AK_HOME = 0, AK_WORK = 1, AK_TRANSPORT = 2, AK_HOUSING = 3, AK_HOSPITAL0 = 4, AK_HOSPITAL1 = 5,  // Citizen.current_area
};
constexpr uint32_t ST_STATE_MASK = 0x7u;
constexpr int ST_SEV_SHIFT = 3;        // 2 bits
constexpr int ST_IMM_SHIFT = 5;        // 3 bits, immunity + 2
constexpr uint32_t ST_VACC = 1u << 8;  // vaccinated
constexpr uint32_t ST_PT = 1u << 9;    // uses_public_transport
constexpr uint32_t ST_HOSP = 1u << 10; // hospitalized
constexpr uint32_t ST_ISO = 1u << 11;  // isolated
constexpr uint32_t ST_WQ = 1u << 12;   // work_quarantined
constexpr int ST_WS_SHIFT = 13;        // 2 bits
constexpr int ST_AREA_SHIFT = 15;      // 3 bits
constexpr uint32_t ST_AREA_MASK = 7u << ST_AREA_SHIFT;
constexpr int ST_DAY_SHIFT = 18;       // 14 bits infection_day
constexpr uint32_t ST_DAY_MAX = 0x3FFFu;

// ---- cell packing ---------------------------------------------------------------------------------------------------
constexpr int CELL_BITS = 14;
constexpr uint32_t CELL_XMASK = (1u << CELL_BITS) - 1;
constexpr uint32_t MAX_COORD = CELL_XMASK;  // grid_size <= 16382

// ---- proposal word ----------------------------------------------------------------------------------------------------
constexpr uint32_t PROP_CELL_MASK = (1u << 28) - 1;
constexpr uint32_t PROP_MOVE = 1u << 28;   // agent wants to move to bits 0..27
constexpr uint32_t PROP_DIRTY = 1u << 29;  // the agent's grid byte changed
constexpr int PROP_BYTE_SHIFT = 30;        // new grid byte - 1

constexpr uint32_t CELL_OCC_MASK = 0x3u;
constexpr uint32_t HOSP_NONE = 0xFFFFFFFFu;
constexpr uint32_t TOT_COPIES = 32;  // power of two, <= 32
constexpr uint32_t ORIGIN_MASK = (1u << 28) - 1;  // home / work words: packed origin in the low 28 bits
constexpr uint32_t GRID_XOFF = 16;                // zero bytes before the start of row 0 in the grid allocation
constexpr uint32_t GRID_YOFF = 3;                 // zero rows above row 0 and below the last row

struct Rect {
    int sx, sy, ex, ey;  // inclusive on both ends (geography/area.rs:83-88)
};

// run constants, passed by value to every kernel
struct Params {
    uint32_t n;        // agent slots (== population for a standalone engine; slots whose state is ST_ABSENT are empty)
    int grid_size;     // G: is_point_in_grid is 0 <= x,y < G (allocation_map.rs:156-159)
    uint32_t pitch;    // bytes per grid row (multiple of 16, >= grid_size + 2)
    uint32_t rows;     // grid_size + 1: Area ends are inclusive
    // the shared rectangles a current_area kind >= AK_TRANSPORT names: zone[kind - AK_TRANSPORT] =
    // transport strip, housing strip (geography/mod.rs:33-70), hospital after Grid::resize_hospital, hospital after
    // increase_hospital_size (grid.rs:233-261)
    Rect zone[4];
    Rect work;         // the work strip (offices live here)
    int hospital_gen;  // which hospital rectangle is grid.hospital_area now: zone[2 + hospital_gen]
    int house_nx, office_nx;        // houses / offices per row (area_factory, geography/area.rs:95-117)
    int house_ny, office_ny;        // rows of houses / offices
    // disease (common/src/disease/mod.rs:26-45)
    uint32_t regular_start, high_start, last_day;
    uint32_t exposed_duration, pre_symptomatic_duration;
    uint64_t thr_rate[3];  // Bernoulli thresholds for rate class 0 (never), 1 (regular), 2 (high)
    uint64_t thr_death, thr_symptomatic, thr_severe;
    uint32_t hospitalize_mask;  // bit c set: rate class c satisfies Disease::is_to_be_hospitalized (disease/mod.rs:97-99)
    uint32_t id_bits;           // claim word: low id_bits = id_mask - agent
    uint64_t seed;
    uint32_t rk[10][2];         // Philox round keys of `seed` (key schedule hoisted out of the kernels)
    int region;
    // byte offset of cell (x, y) in grid[], and its index in claim[]
    // (32-bit arithmetic: rows x pitch <= 16390 x 16400 < 2^32, and x + GRID_XOFF >= 0 for every cell a window can touch)
    __host__ __device__ uint32_t cell_index(int x, int y) const { return (uint32_t)(y + (int)GRID_YOFF) * pitch + (uint32_t)(x + (int)GRID_XOFF); }
    __host__ __device__ size_t cell_offset(int x, int y) const { return cell_index(x, y); }
    __host__ __device__ size_t cell_offset(uint32_t packed) const { return cell_offset((int)(packed & 0x3FFFu), (int)(packed >> 14)); }
    __host__ __device__ size_t grid_bytes() const { return (size_t)pitch * (rows + 2u * GRID_YOFF) + 2u * GRID_XOFF; }
    __host__ __device__ const Rect& transport() const { return zone[0]; }
    __host__ __device__ const Rect& housing() const { return zone[1]; }
    __host__ __device__ const Rect& hospital() const { return zone[2 + hospital_gen]; }
};

// one traveller on the wire (engine/src/travel/commute/commuter.rs:26-35, migration/migrator.rs:25-32): 8 words
struct TravelRecord {
    uint32_t st;    // packed state word of the sender (state, severity, day, immunity, vaccinated, uses_public_transport, work status)
    uint32_t t0;    // at_hour of Exposed / Pre
    uint32_t home;  // Commuter.home_location origin (meaningless for a Migrator: the receiver assigns a house)
    uint32_t work;  // Commuter.work_location origin
    uint32_t reg;   // home region | work region << 8
    uint32_t slot;  // the sender's slot (informational, like the reference's Uuid)
    uint32_t from;  // sending region
    uint32_t pad;
};

enum : int { TRAVEL_MIGRATE = 0, TRAVEL_COMMUTE = 1 };
struct TravelArgs {
    int kind;               // TRAVEL_MIGRATE (h % 24 == 0) or TRAVEL_COMMUTE (h % 24 in {7, 17})
    uint32_t hour, hour_of_day;
    uint64_t thr_outgoing;  // Bernoulli threshold of EngineMigrationPlan::percent_outgoing (filled in on the device: it depends on the current population)
    uint32_t row_index;     // counts ring row that receives the Counts of the exchange hour (k_travel_place), or 0xFFFFFFFF
};

// ---- traveller exchange: device-resident bookkeeping (travel.cu) ------------------------------------------------------------
constexpr uint32_t TRAVEL_MAX_REGIONS = 256;
constexpr uint32_t OCC_ABSENT = 0xFFFFFFFFu;  // a house / office that is not in the occupancy heap (grid.rs:125-155: houses without residents)
constexpr uint32_t HOUSE_CAP = 4, OFFICE_CAP = 100;  // HOME_SIZE^2, OFFICE_SIZE^2 (constants.rs:45-46)
enum : uint32_t {
    TERR_LIST_OVERFLOW = 1, TERR_SEGMENT_OVERFLOW = 2, TERR_NO_HOUSE = 4, TERR_NO_OFFICE = 8, TERR_HOUSES_FULL = 16, TERR_OFFICES_FULL = 32,
    TERR_BAD_REGION = 64, TERR_NO_SLOTS = 128, TERR_PERCENT = 256, TERR_NO_PLACE = 512, TERR_PEER_TIMEOUT = 1024,
};
// scalars of the exchange; the host reads them when it collects Counts rows (errors are sticky)
struct TravelVars {
    uint32_t total;      // leaving candidates, in ascending slot order
    uint32_t n_send;     // records written to the send buffer
    uint32_t n_working;  // arriving migrators that take an office
    uint32_t pending;    // arrivals still without a cell after the placement rounds so far
    uint32_t err;        // TERR_* flags
    uint32_t n_in;       // arrivals of the exchange being unpacked
    uint32_t free_top;   // height of the free-slot stack
    uint32_t abort;      // err as k_travel_plan saw it: non-zero = this pack removes nobody
    uint32_t cnt[TRAVEL_MAX_REGIONS];   // records per destination region
    uint32_t base[TRAVEL_MAX_REGIONS];  // exclusive prefix of cnt
    uint32_t hist[TRAVEL_MAX_REGIONS];  // commuters per destination while k_travel_leave counts them (zero between exchanges)
    uint32_t population;  // live agents of the region (leavers subtracted by k_travel_leave, arrivals added by k_travel_arrive)
    uint32_t rounds;      // most placement rounds one exchange needed so far (test hook)
    uint32_t pend2[2];    // arrivals still without a cell after placement round a: pend2[a & 1]
};
// the pop sequence of an occupancy heap for K arrivals ("water filling"), see travel.cu
struct FillPlan {
    uint32_t K, L, l_last, last_count;  // K pops; lowest non-empty level; level of the last pop; pops on that last level
    uint32_t start[OFFICE_CAP + 1];     // start[l] = index of the first pop that is served from level l
};
struct TravelPtrs {
    uint32_t *occ_house, *occ_office;  // occupants per house / office in heap tie-break order (rank), OCC_ABSENT = not in the heap
    uint32_t* free_stack;              // free agent slots, LIFO: arrivals pop from the top, departures push
    TravelVars* tv;
    const uint32_t* plan_row;          // this region's row of the migration matrix [n_regions]
    uint32_t *list_slot, *list_dest, *list_pos;  // leaving candidates (ascending slot), their destination, their index inside the destination's segment
    TravelRecord* arrivals;            // received records, contiguous, ordered by source region
    uint32_t *arr_widx, *arr_house, *arr_office;  // per arrival: rank among working arrivals, assigned house / office (rank order index)
    uint8_t* placed;
    uint32_t *table_keys, *table_vals;  // placement round hash table
    uint32_t table_mask;
    uint32_t list_cap;
    uint32_t *bh_house, *pref_house, *bh_office, *pref_office;  // per-block occupancy histograms and per-level block prefixes
    FillPlan *plan_house, *plan_office;
    int n_regions;
    const uint32_t* seg_cap;  // records (header included) the segment of each destination may hold, or nullptr = the stride
    uint32_t* chunk_dest;     // [list_cap / 256 + 1][n_regions] commuters per destination in every chunk of 256 list entries
    uint32_t* hslice;         // water filling scratch, per grid block of k_travel_arrive: [2][OFFICE_CAP][grid] slice totals
    uint32_t *tot_house, *tot_office;  // areas per occupancy level overall, TOT_COPIES spread copies [copy][HOUSE_CAP] / [copy][OFFICE_CAP], kept current with bh_*
    uint32_t* foreign;        // one bit per slot: the agent has its home or its work in another region (kept current by install / leave)
};

// ---- tile kernels of the plain movement hours (tiles.cu) ------------------------------------------------------------------------
// A tile = up to tile_units x-adjacent units (10x10 offices / 2x2 houses) of one unit row; tile index = unit row * chunks + chunk.
struct TileGeom {
    int ox, oy;            // origin of unit (0, 0): the work strip's / housing strip's top-left cell
    int unit;              // cells per unit side: OFFICE_SIZE = 10, HOME_SIZE = 2 (constants.rs:45-46)
    int units_x, units_y;  // units per row, rows (area_factory, geography/area.rs:95-117)
    int tile_units, chunks;  // units per tile along x; tiles per unit row
    int cls;               // RC_OFFICE / RC_HOME (hour_common.cuh)
    int sp;                // bytes per staged grid row in shared memory
    uint32_t n_tiles, cap; // tiles; members a tile can settle on chip (more: the tile takes the global path)
    uint32_t threads;      // CTA size of the tile kernel
};
struct TilePtrs {
    uint32_t* perm;             // agent slots: tile 0's local members (ascending id), its riders, tile 1's, ..., then the generic segment
    uint32_t* start;            // [2 * n_tiles + 2] start[2t] = tile t's local members, start[2t + 1] = its riders, start[2 * n_tiles] = the generic segment, then = slots
    uint32_t* dirty;            // [n_tiles] set by the generic segment: somebody outside the tile proposes into it this hour
    const uint32_t* n_housing;  // agents whose current_area is the housing strip (house tiles are off while it is non-zero)
    uint32_t* misc;             // [0] = n_housing, [1] = "a tile order is stale" flag
};

struct Clock {           // device-resident so CUDA graphs can be replayed for any day
    uint32_t hour_base;  // kernels run hour = hour_base + offset
    uint32_t epoch_base; // claim stamp = hour - epoch_base + 1
    uint32_t ring_base;  // Counts of `hour` accumulate in counts ring row hour - ring_base
    uint32_t pad;
};

struct DevPtrs {
    uint32_t *cell, *st, *t0, *home, *work, *wsa, *prop;
    uint32_t* reg;         // home region | work region << 8 (Area.location_id of home_location / work_location); travel kernels only
    uint8_t* grid;
    uint32_t* claim;
    uint32_t* counts;      // ring [rows][8]: S,E,I,H,R,D,pad,pad
    uint32_t* tot;         // running totals [TOT_COPIES][8]; the Counts of the region are the column sums (mod 2^32)
    uint32_t* hosp_first;  // rank of the first vacant hospital cell in Area::iter order, or HOSP_NONE
    const Clock* clock;
    const uint64_t* draws; // injected draws table or nullptr
    unsigned long long* trace;  // debugging (EPI_TRACE=1): globaltimer stamps, [0] = next free entry; entries = (tag << 56) | nanoseconds; or nullptr
};

}  // namespace epi
