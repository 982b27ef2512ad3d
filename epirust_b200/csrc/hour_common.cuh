// Device code shared by the hour kernels (kernels.cu: one agent per thread in id order) and the tile kernels (tiles.cu: one CTA
// per office / house tile with the tile's grid bytes and claim words in shared memory): grid-window helpers, the Philox draw
// schedule and agent_hour(), the whole agent-hour of the reference (citizen/mod.rs:227-432, default_disease_handler.rs:31-103,
// counts.rs:126-140) against the grid in global memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "agent.cuh"
#include "layout.h"
#include "philox.cuh"

namespace epi {

#ifndef EPI_HBS
#define EPI_HBS 128  // threads per CTA of k_hour: 12 CTAs / SM (32: 0.215 ms, 64: 0.193, 128: 0.189, 256: 0.190, 512: 0.247 per launch at 10 M agents)
#endif
#ifndef EPI_MINB
#define EPI_MINB 6  // resident CTAs per SM the movement-hour kernel is compiled for (40 registers; 8 -> 32 registers + spills, measured slower)
#endif
enum : int { MODE_STAY = 0, MODE_WALK = 1, MODE_GOTO = 2 };
enum : int { KIND_START = 0, KIND_MOVE = 1, KIND_END = 2 };

// Loads the compiler may not sink below a branch: all of an agent's words are requested in one memory round trip.  The
// per-agent arrays are streamed once per kernel, so they carry the evict-first hint (.cs) and leave the L2 to the grid and
// the claim words, which are the randomly accessed data.
__device__ __forceinline__ uint32_t ld_early(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cs.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_early_rw(const uint32_t* p) {  // for arrays this kernel also writes
    uint32_t v;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_plain(const uint32_t* p) {  // the tile kernels: neighbouring warps come back to the same sectors
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// Every thread asks the L2 for the per-agent words the thread one wave of CTAs ahead will load first (148 SMs x 6 CTAs x 256
// agents): the first of the two dependent memory round trips of an agent-hour then costs an L2 hit instead of a DRAM access
// (+5 % agent-steps/s at 10 M agents; half a wave is as good, 2 and 4 waves are worse; prefetching the grid rows of the agent
// ahead as well costs more issue slots than it saves).
constexpr uint32_t PREFETCH_AHEAD = 148u * 6u * 256u;  // one wave of k_hour: 148 SMs x 48 warps
constexpr uint32_t PREFETCH_MIN_AGENTS = 4u << 20;       // below this the per-agent arrays (32 B per agent) and the grids fit the 126 MB L2
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ bool rect_contains(const Rect& r, int x, int y) { return r.sx <= x && r.ex >= x && r.sy <= y && r.ey >= y; }

__device__ __forceinline__ Rect origin_rect(uint32_t packed, int size_minus_1) {
    Rect r;
    r.sx = (int)(packed & CELL_XMASK);
    r.sy = (int)((packed >> CELL_BITS) & CELL_XMASK);
    r.ex = r.sx + size_minus_1;
    r.ey = r.sy + size_minus_1;
    return r;
}

// The 8 Moore neighbours of a cell in the reference's iterator order (geography/point.rs:59):
// j: 0 (-1,-1) 1 (0,-1) 2 (1,-1) 3 (-1,0) 4 (1,0) 5 (-1,1) 6 (0,1) 7 (1,1)
struct Hood {
    uint32_t lo, hi;  // grid bytes of neighbours 0..3 and 4..7
};
// 5x5 window of grid bytes centred on (cx, cy).  Row k (dy = k - 2): l[k] holds cells cx-2..cx+1, r[k] cells cx-1..cx+2.
struct Window {
    uint32_t l[5], r[5];
};
__device__ __forceinline__ Window load_window(const uint8_t* __restrict__ grid, const Params& P, int cx, int cy) {
    const uint32_t first = P.cell_index(cx - 2, cy - 2);  // the window's top-left cell; GRID_XOFF and pitch are multiples of 4
    const uint32_t sh = (first & 3u) * 8u;
    const uint32_t* g32 = reinterpret_cast<const uint32_t*>(grid);
    const uint32_t w0 = first >> 2, stride = P.pitch >> 2;  // 32-bit word indices: one widening multiply-add per row address
    uint32_t a[5], b[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t* p = g32 + (w0 + (uint32_t)k * stride);
        a[k] = __ldg(p);
        b[k] = __ldg(p + 1);
    }
    Window win;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        win.l[k] = __funnelshift_r(a[k], b[k], sh);
        win.r[k] = __funnelshift_rc(a[k], b[k], sh + 8u);
    }
    return win;
}
// the 3x3 neighbourhood of the window's centre cell
__device__ __forceinline__ Hood hood_centre(const Window& w) {
    Hood h;
    h.lo = (w.r[1] & 0x00FFFFFFu) | (w.r[2] << 24);
    h.hi = ((w.r[2] >> 16) & 0xFFu) | (w.r[3] << 8);
    return h;
}
// the 3x3 neighbourhood of the cell at offset (dx, dy) from the window centre, dx, dy in {-1, 0, 1}
__device__ __forceinline__ Hood hood_at(const Window& w, int dx, int dy) {
    uint32_t t0 = dx < 0 ? w.l[0] : w.r[0], t1 = dx < 0 ? w.l[1] : w.r[1], t2 = dx < 0 ? w.l[2] : w.r[2];
    const uint32_t t3 = dx < 0 ? w.l[3] : w.r[3], t4 = dx < 0 ? w.l[4] : w.r[4];
    if (dy == 0) { t0 = t1; t1 = t2; t2 = t3; }
    else if (dy > 0) { t0 = t2; t1 = t3; t2 = t4; }
    const uint32_t sh = dx > 0 ? 8u : 0u;
    const uint32_t top = t0 >> sh, mid = t1 >> sh, bot = t2 >> sh;
    Hood h;
    h.lo = (top & 0x00FFFFFFu) | (mid << 24);
    h.hi = ((mid >> 16) & 0xFFu) | (bot << 8);
    return h;
}
// one bit per byte (the 0x01 position of each byte of `bits`) -> 4-bit mask
__device__ __forceinline__ uint32_t gather4(uint32_t bits) { return ((bits & 0x01010101u) * 0x01020408u) >> 24; }
// bit j set: neighbour j's cell is vacant (occupancy bits 0-1 clear; claim bits ignored)
__device__ __forceinline__ uint32_t vacant_mask(const Hood& h) {
    return gather4(~(h.lo | (h.lo >> 1))) | (gather4(~(h.hi | (h.hi >> 1))) << 4);
}
// bit j set: neighbour j holds an infected, not hospitalized agent with a non-zero rate class (occupancy value 2 or 3)
__device__ __forceinline__ uint32_t infectious_mask(const Hood& h) { return gather4(h.lo >> 1) | (gather4(h.hi >> 1) << 4); }

// Area::get_neighbors_of(c).filter(is_point_in_grid): which neighbours of (cx, cy) lie inside rectangle r and the grid
// (geography/area.rs:56-58, allocation_map.rs:156-159).  (cx, cy) itself need not be inside r.
__device__ __forceinline__ uint32_t valid_mask(const Rect& r, int G, int cx, int cy) {
    const int ex = min(r.ex, G - 1), ey = min(r.ey, G - 1);  // r.sx, r.sy >= 0 always
    const uint32_t cl = (cx - 1 >= r.sx) & (cx - 1 <= ex), cc = (cx >= r.sx) & (cx <= ex), cr = (cx + 1 >= r.sx) & (cx + 1 <= ex);
    const uint32_t ru = (cy - 1 >= r.sy) & (cy - 1 <= ey), rc = (cy >= r.sy) & (cy <= ey), rd = (cy + 1 >= r.sy) & (cy + 1 <= ey);
    const uint32_t cols = cl | (cc << 1) | (cr << 2);
    return (cols & (0u - ru)) | (((cl | (cr << 1)) & (0u - rc)) << 3) | ((cols & (0u - rd)) << 5);
}
// the same when (cx, cy) is known to lie inside r
__device__ __forceinline__ uint32_t valid_mask_inside(const Rect& r, int G, int cx, int cy) {
    const uint32_t cl = cx > r.sx, cc = cx < G, cr = cx < min(r.ex, G - 1);
    const uint32_t ru = cy > r.sy, rc = cy < G, rd = cy < min(r.ey, G - 1);
    const uint32_t cols = cl | (cc << 1) | (cr << 2);
    return (cols & (0u - ru)) | (((cl | (cr << 1)) & (0u - rc)) << 3) | ((cols & (0u - rd)) << 5);
}
// position of the idx-th set bit of an 8-bit mask (idx < popc(m))
__device__ __forceinline__ int select_bit(uint32_t m, uint32_t idx) {
    int base = 0;
    uint32_t c = __popc(m & 0xFu);
    if (idx >= c) { idx -= c; m >>= 4; base = 4; }
    c = __popc(m & 0x3u);
    if (idx >= c) { idx -= c; m >>= 2; base += 2; }
    c = m & 1u;
    if (idx >= c) base += 1;
    return base;
}
__device__ __forceinline__ int hood_dx(int j) { return (int)((0x9224u >> (2 * j)) & 3u) - 1; }  // {-1,0,1,-1,1,-1,0,1}
__device__ __forceinline__ int hood_dy(int j) { return (int)((0xA940u >> (2 * j)) & 3u) - 1; }  // {-1,-1,-1,0,0,1,1,1}

// Philox4x32-10 with the key schedule taken from Params (constant bank operands)
__device__ __forceinline__ U4 philox_rk(const Params& P, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ P.rk[r][0];
        c2 = hi0 ^ c3 ^ P.rk[r][1];
        c1 = lo1;
        c3 = lo0;
    }
    return U4{c0, c1, c2, c3};
}

template <bool INJECT>
struct Draws {
    const Params& P;
    uint32_t agent, hour;
    const uint64_t* row;
    __device__ __forceinline__ U4 block(uint32_t b) const { return philox_rk(P, agent, hour, b, DOM_STEP); }
    // block 0: PICK, FACTOR, A
    __device__ __forceinline__ void common(uint32_t& pick, uint32_t& factor, uint64_t& a) const {
        if (INJECT) { pick = (uint32_t)row[SLOT_PICK]; factor = (uint32_t)row[SLOT_FACTOR]; a = row[SLOT_A]; return; }
        const U4 o = block(0);
        pick = o.x; factor = o.y; a = u64_of(o.z, o.w);
    }
    // block 1: Area::get_random_point (geography/area.rs:76-81)
    __device__ __forceinline__ void point(const Rect& r, int& px, int& py) const {
        uint32_t dx, dy;
        if (INJECT) { dx = (uint32_t)row[SLOT_PX]; dy = (uint32_t)row[SLOT_PY]; }
        else { const U4 o = block(1); dx = o.x; dy = o.y; }
        px = r.sx + (int)__umulhi(dx, (uint32_t)(r.ex - r.sx + 1));
        py = r.sy + (int)__umulhi(dy, (uint32_t)(r.ey - r.sy + 1));
    }
    __device__ __forceinline__ uint64_t expose(int j) const {
        if (INJECT) return row[SLOT_EXPOSE0 + j];
        const U4 o = block(2u + ((uint32_t)j >> 1));
        return (j & 1) ? u64_of(o.z, o.w) : u64_of(o.x, o.y);
    }
};

// hour_base and epoch_base of the device-resident clock in one 64-bit load
__device__ __forceinline__ uint2 load_clock(const Clock* c) { return *reinterpret_cast<const uint2*>(c); }

// Where an agent-hour reads its grid windows and what becomes of its proposal.  GlobalEnv: the grid in global memory, the
// proposal goes to prop[] + an atomicMax on claim[] for k_commit (kernels.cu).  tiles.cu has the shared-memory environment.
enum : int { RC_HOME = 0, RC_OFFICE = 1, RC_ZONE = 2 };  // what kind of rectangle an agent's movement is confined to this hour
template <bool ALWAYS_WRITE_PROP, bool STREAM = true>
struct GlobalEnv {
    static constexpr bool stream_loads = STREAM;  // the per-agent words are read once per kernel: evict-first
    const Params& P;
    const DevPtrs& D;
    uint32_t epoch_base;  // Clock::epoch_base, read together with hour_base by the caller (load_clock): one load instead of two
    __device__ __forceinline__ void on_rule(int, const Rect&, int) {}
    __device__ __forceinline__ Window window(int cx, int cy) const { return load_window(D.grid, P, cx, cy); }
    // the agent stands on (x, y), proposes (tx, ty); `dirty`: its grid byte changes; byte = its new grid byte
    __device__ __forceinline__ void commit(uint32_t i, uint32_t hour, int x, int y, int tx, int ty, bool dirty, uint32_t byte) {
        uint32_t prop = dirty ? PROP_DIRTY : 0u;
        const uint32_t target = ((uint32_t)ty << CELL_BITS) | (uint32_t)tx;
        if (target != (((uint32_t)y << CELL_BITS) | (uint32_t)x)) {
            prop |= PROP_MOVE | target;
            const uint32_t stamp = hour - epoch_base + 1u;
            const uint32_t id_mask = (1u << P.id_bits) - 1u;
            atomicMax(&D.claim[P.cell_index(tx, ty)], (stamp << P.id_bits) | (id_mask - i));
        }
        if (prop) prop |= (byte - 1u) << PROP_BYTE_SHIFT;
        if (ALWAYS_WRITE_PROP || prop) st_stream(D.prop + i, prop);
    }
};

// The agent-hour of agent slot i.  KIND: the hour-of-day class (ROUTINE_START_TIME, ROUTINE_END_TIME, else perform_movements);
// HOD: the hour of day the movement code is compiled for.  perform_movements (citizen/mod.rs:257-349) has special cases at
// h = 7, 8, 12, 16, 17 only; each of them and "any other hour" (HOD = 9: everybody who can move walks inside current_area,
// 11 of the 16 movement hours) gets its own instantiation, so the goto / area-change logic of the other hours is compiled
// out (the kernel is issue-and-latency bound: -10 % time on the plain hours against one kernel with a run-time hour).
template <int KIND, bool INJECT, uint32_t HOD, class Env>
__device__ __forceinline__ void agent_hour(const Params& P, const DevPtrs& D, uint32_t i, uint32_t hour, Env& env) {
    constexpr uint32_t h = HOD;
    // one round trip: the agent's state words (and the uniform clock word)
    const uint32_t s0 = Env::stream_loads ? ld_early_rw(D.st + i) : ld_plain(D.st + i);
    const uint32_t c0 = Env::stream_loads ? ld_early_rw(D.cell + i) : ld_plain(D.cell + i);
    const uint32_t hm = Env::stream_loads ? ld_early(D.home + i) : ld_plain(D.home + i);
    const uint32_t wk = KIND == KIND_MOVE ? (Env::stream_loads ? ld_early(D.work + i) : ld_plain(D.work + i)) : 0u;
    if ((s0 & ST_STATE_MASK) == ST_ABSENT) return;  // empty slot; its prop word stays 0
    const int x = (int)(c0 & CELL_XMASK), y = (int)(c0 >> CELL_BITS);
    const uint8_t* __restrict__ grid = D.grid;
    uint32_t s = s0;
    int tx = x, ty = y;
    const Draws<INJECT> dr{P, i, hour, INJECT ? D.draws + (size_t)i * 16 : nullptr};
    const uint32_t ws = (s0 >> ST_WS_SHIFT) & 3u;
    uint32_t state = s0 & ST_STATE_MASK, sev = (s0 >> ST_SEV_SHIFT) & 3u, day = s0 >> ST_DAY_SHIFT;
    bool transition = KIND != KIND_MOVE;  // state / sev / day may differ from s0: the state word is re-packed at the end
    const Rect home = origin_rect(hm, 1);

    if (KIND == KIND_START) {
        // ROUTINE_START_TIME: increment_infection_day + hospitalize (citizen/mod.rs:240-243, :351-365)
        if (state == ST_I) {
            day = min(day + 1u, ST_DAY_MAX);
            const int imm = (int)((s0 >> ST_IMM_SHIFT) & 7u) - 2;
            if (!(s0 & ST_HOSP) && sev == SEV_SEVERE && ((P.hospitalize_mask >> rate_class(P, (uint32_t)((int)day + imm))) & 1u)) {
                const uint32_t first = *D.hosp_first;
                if (first != HOSP_NONE) {  // goto_hospital: every admitted agent targets the same first vacant cell
                    const Rect hr = P.hospital();
                    const uint32_t w = (uint32_t)(hr.ex - hr.sx + 1);
                    tx = hr.sx + (int)(first % w); ty = hr.sy + (int)(first / w);
                    s |= ST_HOSP;
                } else {  // hospital full: try a random point of the own house
                    int px, py;
                    dr.point(home, px, py);
                    if ((grid[P.cell_offset(px, py)] & CELL_OCC_MASK) == 0) { tx = px; ty = py; }
                }
            }
        }
    } else if (KIND == KIND_END) {
        // ROUTINE_END_TIME: Citizen::deceased + on_routine_end (citizen/mod.rs:397-413, default_disease_handler.rs:88-103)
        if (state == ST_I) {
            if ((sev == SEV_ASYM && day == 9u) || (sev == SEV_MILD && day == 12u)) state = ST_R;
            else if (sev == SEV_SEVERE && day == P.last_day) {
                uint32_t pick, factor; uint64_t a;
                dr.common(pick, factor, a);
                state = bernoulli(a, P.thr_death) ? ST_D : ST_R;
            }
        }
        if (state == ST_R) {  // every recovered agent, every day
            int px, py;
            dr.point(home, px, py);
            if ((grid[P.cell_offset(px, py)] & CELL_OCC_MASK) == 0) { tx = px; ty = py; }
        }
        if (state == ST_R || state == ST_D) { s &= ~ST_HOSP; sev = 0; day = 0; }
    } else {
        // perform_movements (citizen/mod.rs:257-349); h in 7..22 here (sleep hours use k_sleep)
        const uint32_t kind0 = (s0 >> ST_AREA_SHIFT) & 7u;
        const bool symptomatic = state == ST_I && sev >= SEV_MILD;
        const bool can_move = !(symptomatic || (s0 & (ST_HOSP | ST_ISO)) || state == ST_D);  // citizen/mod.rs:452-454
        const bool pre = state == ST_I && sev == SEV_PRE;
        // at_hour of Exposed / Pre: requested now, consumed after the window arrives
        uint32_t t0v = 0;
        if (state == ST_E || pre) t0v = ld_early_rw(D.t0 + i);
        // Citizen.work_location of a non-working agent is its home (citizen_factory.rs:62-64).  The packed origin is chosen first
        // and unpacked once (the ALU pipe is this kernel's limit: every select and shift counts).
        auto rect_of = [&](uint32_t kind) -> Rect {
            const bool office = kind == AK_WORK && ws != WS_NA;
            Rect r = origin_rect(office ? wk : hm, office ? 9 : 1);
            if (kind >= AK_TRANSPORT) r = P.zone[kind - AK_TRANSPORT];
            return r;
        };
        auto work_rect = [&]() -> Rect { return rect_of(AK_WORK); };
        // the hour's rule -> (mode, rectangle R, new current_area kind).  goto_area for a non-working agent is
        // move_agent_from in the (old) current_area (citizen/mod.rs:386-394), i.e. MODE_WALK.
        auto class_of = [&](uint32_t kind) -> int { return kind >= AK_TRANSPORT ? RC_ZONE : (kind == AK_WORK && ws != WS_NA) ? RC_OFFICE : RC_HOME; };
        Rect R = rect_of(kind0);
        int rcls = class_of(kind0);
        uint32_t kind = kind0;
        int mode = MODE_WALK;
        bool dynamics = true, override_movement = false;
        if (ws == WS_NA) {
            if (h == 8) kind = AK_HOUSING;
            else if (h == 12) kind = AK_HOME;
        } else if (ws != WS_STAFF) {  // Normal | Essential
            if (h == 7 || h == 17) {
                if (s0 & ST_PT) { mode = MODE_GOTO; R = P.transport(); kind = AK_TRANSPORT; rcls = RC_ZONE; }
            } else if (h == 8) { mode = MODE_GOTO; R = work_rect(); kind = AK_WORK; rcls = RC_OFFICE; }
            else if (h == 16) {
                mode = MODE_GOTO; R = home; kind = AK_HOME; rcls = RC_HOME;
                override_movement = symptomatic && rect_contains(work_rect(), x, y);  // citizen/mod.rs:373-381
            }
        } else {  // HospitalStaff { work_start_at } (0.14 % of agents)
            const uint32_t wsa = D.wsa[i];
            const uint32_t since = hour >= wsa ? hour - wsa : 0u;  // saturating_sub
            const uint32_t hosp_kind = P.hospital_gen ? AK_HOSPITAL1 : AK_HOSPITAL0;
            if (since == 24u * 14u) { s |= ST_WQ; dynamics = false; mode = MODE_STAY; }
            else if (since == 24u * 14u * 2u) {
                mode = MODE_GOTO; R = home; kind = AK_HOME; rcls = RC_HOME;
                D.wsa[i] = hour + 24u * 14u;
                dynamics = false;
            } else if (h == 8) {
                mode = MODE_STAY;
                if (kind0 != hosp_kind && wsa <= hour) {
                    mode = MODE_GOTO; R = P.hospital(); kind = hosp_kind; rcls = RC_ZONE;
                    D.wsa[i] = hour;
                }
                s &= ~ST_WQ;
            } else if (h == 16) { s |= ST_WQ; mode = MODE_STAY; }
            else if (s0 & ST_WQ) mode = MODE_STAY;
        }
        if (!(can_move || override_movement)) mode = MODE_STAY;
        env.on_rule(rcls, R, mode);
        s = (s & ~ST_AREA_MASK) | (kind << ST_AREA_SHIFT);

        const bool in_area = rect_contains(R, x, y);
        const bool need_point = mode == MODE_GOTO || (mode == MODE_WALK && !in_area);
        const bool scan = dynamics && state == ST_S && !(s & (ST_WQ | ST_VACC));
        int bx = x, by = y;
        if (need_point) dr.point(R, bx, by);
        // second round trip: the 5x5 window around the base cell
        Window win;
        if (mode != MODE_STAY || scan) win = env.window(bx, by);
        uint32_t pick = 0, factor = 0;
        uint64_t a = 0;
        if (mode == MODE_WALK || (dynamics && (state == ST_E || pre))) dr.common(pick, factor, a);
        int ddx = 0, ddy = 0;  // proposed cell relative to the window centre (bx, by)
        bool in_window = !need_point;  // the cell the agent ends up proposing / standing on is (bx + ddx, by + ddy), inside the loaded window
        if (mode == MODE_GOTO) {
            if (((win.r[2] >> 8) & CELL_OCC_MASK) == 0) { tx = bx; ty = by; in_window = true; }  // target.get_random_point vacant -> go (citizen/mod.rs:387-392)
        } else if (mode == MODE_WALK) {  // Citizen::move_agent_from (citizen/mod.rs:415-432)
            const uint32_t cand = vacant_mask(hood_centre(win)) & valid_mask_inside(R, P.grid_size, bx, by);
            if (cand) {  // candidates.choose(rng): the k-th candidate in iterator order, k uniform
                const int j = select_bit(cand, __umulhi(pick, (uint32_t)__popc(cand)));
                ddx = hood_dx(j); ddy = hood_dy(j);
                tx = bx + ddx; ty = by + ddy;
                in_window = true;
            }
        }
        if (scan) {  // on_susceptible at the proposed cell (disease_state_machine.rs:53-70, default_disease_handler.rs:64-86)
            // the proposed cell is the window centre + (ddx, ddy), or the agent's own cell when the move was
            // not possible: a relocated walker / a goto that found its point occupied stays at (x, y), which
            // is outside the window -> second load (hours 7, 8, 16, 17 mostly)
            Hood hd;
            uint32_t inf = 0;
            // nobody infectious anywhere in the 5x5 window (the usual case outside the peak of the epidemic): no neighbourhood to
            // assemble.  Six logic instructions against the ~25 of hood_at + infectious_mask.
            const uint32_t any_inf = (win.l[0] | win.r[0] | win.l[1] | win.r[1] | win.l[2] | win.r[2] | win.l[3] | win.r[3] | win.l[4] | win.r[4]) & 0x02020202u;
            if (!in_window || any_inf) {
                if (in_window) hd = hood_at(win, ddx, ddy);
                else hd = hood_centre(env.window(tx, ty));
                inf = infectious_mask(hd);
            }
            // neighbours are clipped to the NEW current_area: R is its rectangle unless a non-working agent's
            // area changed this hour (h = 8, 12), where R is still the old one
            if (inf) inf &= valid_mask(kind == kind0 ? R : rect_of(kind), P.grid_size, tx, ty);
            while (inf) {
                const int j = __ffs(inf) - 1;
                inf &= inf - 1u;
                const uint32_t b = ((j < 4 ? hd.lo : hd.hi) >> (8 * (j & 3))) & CELL_OCC_MASK;
                if (bernoulli(dr.expose(j), P.thr_rate[b - 1u])) {
                    state = ST_E;
                    transition = true;
                    D.t0[i] = hour;
                    break;
                }
            }
        }
        if (dynamics) {
            // DiseaseStateMachine::next for the other states (disease_state_machine.rs:53-70)
            if (state == ST_E && (s0 & ST_STATE_MASK) == ST_E) {  // on_exposed, :52-62
                const int f = (int)__umulhi(factor, 3u) - 1;
                if (hour - t0v >= (uint32_t)((int)P.exposed_duration + f)) {
                    const bool symptoms = bernoulli(a, P.thr_symptomatic);
                    transition = true;
                    state = ST_I; day = 0;
                    sev = symptoms ? SEV_PRE : SEV_ASYM;
                    if (symptoms) D.t0[i] = hour;
                }
            } else if (pre) {  // on_infected, :41-50
                if (hour - t0v >= P.pre_symptomatic_duration) { sev = bernoulli(a, P.thr_severe) ? SEV_SEVERE : SEV_MILD; transition = true; }
            }
        }
    }
    // re-pack (a movement hour changes state / severity / day only through one of the transitions above)
    if (transition)
        s = (s & ~(ST_STATE_MASK | (3u << ST_SEV_SHIFT) | (ST_DAY_MAX << ST_DAY_SHIFT))) | state | (sev << ST_SEV_SHIFT) | (day << ST_DAY_SHIFT);
    bool dirty = false;
    if (s != s0) {
        st_stream(D.st + i, s);
        // Counts::update_counts (counts.rs:126-140), incrementally: only an agent whose column changed touches the running
        // totals (TOT_COPIES spread copies against same-address contention); k_commit snapshots them into the hour's row.
        const uint32_t cat0 = count_category(s0), cat1 = count_category(s);
        if (cat0 != cat1) {
            uint32_t* t = D.tot + (blockIdx.x & (TOT_COPIES - 1u)) * 8u;
            atomicAdd(t + cat1, 1u);
            atomicSub(t + cat0, 1u);
        }
        dirty = cell_byte(P, s) != cell_byte(P, s0);
    }
    env.commit(i, hour, x, y, tx, ty, dirty, cell_byte(P, s));
}

}  // namespace epi
