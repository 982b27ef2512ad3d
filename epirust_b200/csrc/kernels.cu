// sm_100a kernels of the per-hour agent step.  HBM/L2-bound integer work: no tensor cores.
//
//   k_hospital_scan   first vacant hospital cell in row-major order        (allocation_map.rs:144-147)
//   k_hour<KIND, HOD> one agent per thread: routine, movement proposal against the start-of-hour grid, disease
//                     transition, Counts; atomicMax claim on the target cell (citizen/mod.rs:227-432,
//                     default_disease_handler.rs:31-103, counts.rs:126-140).  KIND: the hour-of-day class
//                     (ROUTINE_START_TIME, ROUTINE_END_TIME, else perform_movements); HOD: the movement hour the kernel is
//                     compiled for (7, 8, 12, 16, 17 or "any other") -- both uniform per launch.
//   k_commit          four agents per thread: lowest-id claimant moves, loser stays; grid bytes updated in place
//                     (allocation_map.rs:93-102,131-134)
//   k_sleep           hours 1..6, four agents per thread: current_area := home (citizen/mod.rs:244-248) + Counts recount
//   k_lock / k_unlock / k_vaccinate   intervention sweeps (allocation_map.rs:349-387)
//
// Synchronous-update argument (why the grid can be updated in place): every proposal targets a cell that was vacant
// at the start of the hour (goto_area / move_agent_from / goto_hospital / deceased all go through
// CitizenLocationMap::move_agent or an is_cell_vacant filter), every cell that is cleared was occupied at the start
// of the hour, so the set of written-to-occupied and written-to-vacant cells are disjoint, and all reads of the grid
// happen in k_hour, all writes in k_commit.
//
// Claim protocol (north_star (b): lowest-agent-id priority): k_hour does atomicMax(claim[cell], stamp | ~id) -- a
// fire-and-forget reduction, nobody waits for its result; k_commit moves the agent iff claim[cell] is its own word.
// The hour stamp in the high bits makes clearing the claim array unnecessary.  Measured alternatives that lost (10 M
// agents, B200): (1) a three-phase variant that marks CLAIMED / CONTESTED bits in the occupancy byte and touches claim[]
// only for contested cells -- its phase-A atomic needs its return value and the extra pass costs more than the DRAM
// traffic it saves (two agents of a house contest a cell every other hour, so "contested" is not rare); (2) atomicMax
// with the old value returned, the displaced claimant notified through a per-agent flag so that k_commit never reads
// claim[] -- k_commit 112 -> 90 us but k_hour 221 -> 258 us; (3) L2 evict_last / evict_first policies on the claim and
// grid accesses by zone -- DRAM bytes unchanged (the 10 M-agent working set does not fit the L2 either way).
//
// Instruction budget: k_hour is issue/latency-bound before it is HBM-bound, so the agent-hour is branch-light: the
// movement rule of the hour is reduced to (mode, rectangle) by predicated integer logic, one Philox block serves the
// common case, the agent's four state words are loaded up front in one round trip, and one 5x5 window of the
// zero-padded grid (10 aligned 32-bit loads, one round trip) serves both the walk and the exposure scan.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "layout.h"
#include "philox.cuh"
#include "agent.cuh"

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
// Every thread asks the L2 for the per-agent words the thread one wave of CTAs ahead will load first (148 SMs x 6 CTAs x 256
// agents): the first of the two dependent memory round trips of an agent-hour then costs an L2 hit instead of a DRAM access
// (+5 % agent-steps/s at 10 M agents; half a wave is as good, 2 and 4 waves are worse; prefetching the grid rows of the agent
// ahead as well costs more issue slots than it saves).
constexpr uint32_t PREFETCH_AHEAD = 148u * 6u * 256u;  // one wave of k_hour: 148 SMs x 48 warps
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
    const size_t first = P.cell_offset(cx - 2, cy - 2);  // the window's top-left cell; GRID_XOFF and pitch are multiples of 4
    const uint32_t sh = (uint32_t)(first & 3u) * 8u;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(grid + (first & ~(size_t)3));
    const uint32_t stride = P.pitch >> 2;
    uint32_t a[5], b[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        a[k] = __ldg(p + (size_t)k * stride);
        b[k] = __ldg(p + (size_t)k * stride + 1);
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

__device__ __forceinline__ void block_count(uint32_t cat, uint32_t* __restrict__ out_row) {
    // warp ballots -> shared -> one atomic per category per block (counts.rs:126-140)
    __shared__ uint32_t s_cnt[6];
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (uint32_t c = 0; c < 6; ++c) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, cat == c);
        if (lane == c && b) atomicAdd(&s_cnt[c], (uint32_t)__popc(b));
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&out_row[threadIdx.x], s_cnt[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_hospital_scan(Params P, const uint8_t* __restrict__ grid, uint32_t* __restrict__ hosp_first) {
    const Rect h = P.hospital();
    const uint32_t w = (uint32_t)(h.ex - h.sx + 1), nh = (uint32_t)(h.ey - h.sy + 1);
    const uint64_t total = (uint64_t)w * nh;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (uint64_t)gridDim.x * blockDim.x) {
        if (r >= *(volatile uint32_t*)hosp_first) return;  // ranks only grow along the stride: nothing better ahead
        const uint32_t x = (uint32_t)h.sx + (uint32_t)(r % w), y = (uint32_t)h.sy + (uint32_t)(r / w);
        if ((grid[P.cell_offset((int)x, (int)y)] & CELL_OCC_MASK) == 0) { atomicMin(hosp_first, (uint32_t)r); return; }
    }
}

// HOD: the hour of day the movement kernel is compiled for.  perform_movements (citizen/mod.rs:257-349) has special cases at
// h = 7, 8, 12, 16, 17 only; each of them and "any other hour" (HOD = 9: everybody who can move walks inside current_area,
// 11 of the 16 movement hours) gets its own instantiation, so the goto / area-change logic of the other hours is compiled
// out (the kernel is issue-and-latency bound: -10 % time on the plain hours against one kernel with a run-time hour).
template <int KIND, bool INJECT, uint32_t HOD = 0>
__global__ void __launch_bounds__(EPI_HBS, KIND == KIND_MOVE ? EPI_MINB * (256 / EPI_HBS) : 4 * (256 / EPI_HBS)) k_hour(Params P, DevPtrs D, uint32_t hour_offset) {
    constexpr uint32_t h = HOD;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    if ((threadIdx.x & 7u) == 0 && i + PREFETCH_AHEAD < P.n) {  // one prefetch per 32-byte sector
        prefetch_l2(D.st + i + PREFETCH_AHEAD);
        prefetch_l2(D.cell + i + PREFETCH_AHEAD);
        prefetch_l2(D.home + i + PREFETCH_AHEAD);
        if (KIND == KIND_MOVE) prefetch_l2(D.work + i + PREFETCH_AHEAD);
    }
    // one round trip: the agent's state words (and the uniform clock word)
    const uint32_t s0 = ld_early_rw(D.st + i);
    const uint32_t c0 = ld_early_rw(D.cell + i);
    const uint32_t hm = ld_early(D.home + i);
    const uint32_t wk = KIND == KIND_MOVE ? ld_early(D.work + i) : 0u;
    const uint32_t hour = D.clock->hour_base + hour_offset;
    if ((s0 & ST_STATE_MASK) == ST_ABSENT) return;  // empty slot; its prop word stays 0
    const int x = (int)(c0 & CELL_XMASK), y = (int)(c0 >> CELL_BITS);
    const uint8_t* __restrict__ grid = D.grid;
    uint32_t s = s0;
    int tx = x, ty = y;
    const Draws<INJECT> dr{P, i, hour, INJECT ? D.draws + (size_t)i * 16 : nullptr};
    const uint32_t ws = (s0 >> ST_WS_SHIFT) & 3u;
    uint32_t state = s0 & ST_STATE_MASK, sev = (s0 >> ST_SEV_SHIFT) & 3u, day = s0 >> ST_DAY_SHIFT;
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
        const Rect workr = ws == WS_NA ? home : origin_rect(wk, 9);
        auto rect_of = [&](uint32_t kind) -> Rect {
            Rect r = kind == AK_WORK ? workr : home;
            if (kind >= AK_TRANSPORT) r = P.zone[kind - AK_TRANSPORT];
            return r;
        };
        // the hour's rule -> (mode, rectangle R, new current_area kind).  goto_area for a non-working agent is
        // move_agent_from in the (old) current_area (citizen/mod.rs:386-394), i.e. MODE_WALK.
        Rect R = rect_of(kind0);
        uint32_t kind = kind0;
        int mode = MODE_WALK;
        bool dynamics = true, override_movement = false;
        if (ws == WS_NA) {
            if (h == 8) kind = AK_HOUSING;
            else if (h == 12) kind = AK_HOME;
        } else if (ws != WS_STAFF) {  // Normal | Essential
            if (h == 7 || h == 17) {
                if (s0 & ST_PT) { mode = MODE_GOTO; R = P.transport(); kind = AK_TRANSPORT; }
            } else if (h == 8) { mode = MODE_GOTO; R = workr; kind = AK_WORK; }
            else if (h == 16) {
                mode = MODE_GOTO; R = home; kind = AK_HOME;
                override_movement = symptomatic && rect_contains(workr, x, y);  // citizen/mod.rs:373-381
            }
        } else {  // HospitalStaff { work_start_at } (0.14 % of agents)
            const uint32_t wsa = D.wsa[i];
            const uint32_t since = hour >= wsa ? hour - wsa : 0u;  // saturating_sub
            const uint32_t hosp_kind = P.hospital_gen ? AK_HOSPITAL1 : AK_HOSPITAL0;
            if (since == 24u * 14u) { s |= ST_WQ; dynamics = false; mode = MODE_STAY; }
            else if (since == 24u * 14u * 2u) {
                mode = MODE_GOTO; R = home; kind = AK_HOME;
                D.wsa[i] = hour + 24u * 14u;
                dynamics = false;
            } else if (h == 8) {
                mode = MODE_STAY;
                if (kind0 != hosp_kind && wsa <= hour) {
                    mode = MODE_GOTO; R = P.hospital(); kind = hosp_kind;
                    D.wsa[i] = hour;
                }
                s &= ~ST_WQ;
            } else if (h == 16) { s |= ST_WQ; mode = MODE_STAY; }
            else if (s0 & ST_WQ) mode = MODE_STAY;
        }
        if (!(can_move || override_movement)) mode = MODE_STAY;
        s = (s & ~ST_AREA_MASK) | (kind << ST_AREA_SHIFT);

        const bool in_area = rect_contains(R, x, y);
        const bool need_point = mode == MODE_GOTO || (mode == MODE_WALK && !in_area);
        const bool scan = dynamics && state == ST_S && !(s & (ST_WQ | ST_VACC));
        int bx = x, by = y;
        if (need_point) dr.point(R, bx, by);
        // second round trip: the 5x5 window around the base cell
        Window win;
        if (mode != MODE_STAY || scan) win = load_window(grid, P, bx, by);
        uint32_t pick = 0, factor = 0;
        uint64_t a = 0;
        if (mode == MODE_WALK || (dynamics && (state == ST_E || pre))) dr.common(pick, factor, a);
        int ddx = 0, ddy = 0;  // proposed cell relative to the window centre (bx, by)
        if (mode == MODE_GOTO) {
            if (((win.r[2] >> 8) & CELL_OCC_MASK) == 0) { tx = bx; ty = by; }  // target.get_random_point vacant -> go (citizen/mod.rs:387-392)
        } else if (mode == MODE_WALK) {  // Citizen::move_agent_from (citizen/mod.rs:415-432)
            const uint32_t cand = vacant_mask(hood_centre(win)) & valid_mask_inside(R, P.grid_size, bx, by);
            if (cand) {  // candidates.choose(rng): the k-th candidate in iterator order, k uniform
                const int j = select_bit(cand, __umulhi(pick, (uint32_t)__popc(cand)));
                ddx = hood_dx(j); ddy = hood_dy(j);
                tx = bx + ddx; ty = by + ddy;
            }
        }
        if (scan) {  // on_susceptible at the proposed cell (disease_state_machine.rs:53-70, default_disease_handler.rs:64-86)
            // the proposed cell is the window centre + (ddx, ddy), or the agent's own cell when the move was
            // not possible: a relocated walker / a goto that found its point occupied stays at (x, y), which
            // is outside the window -> second load (hours 7, 8, 16, 17 mostly)
            Hood hd;
            if (tx == bx + ddx && ty == by + ddy) hd = hood_at(win, ddx, ddy);
            else hd = hood_centre(load_window(grid, P, tx, ty));
            uint32_t inf = infectious_mask(hd);
            // neighbours are clipped to the NEW current_area: R is its rectangle unless a non-working agent's
            // area changed this hour (h = 8, 12), where R is still the old one
            if (inf) inf &= valid_mask(kind == kind0 ? R : rect_of(kind), P.grid_size, tx, ty);
            while (inf) {
                const int j = __ffs(inf) - 1;
                inf &= inf - 1u;
                const uint32_t b = ((j < 4 ? hd.lo : hd.hi) >> (8 * (j & 3))) & CELL_OCC_MASK;
                if (bernoulli(dr.expose(j), P.thr_rate[b - 1u])) {
                    state = ST_E;
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
                    state = ST_I; day = 0;
                    sev = symptoms ? SEV_PRE : SEV_ASYM;
                    if (symptoms) D.t0[i] = hour;
                }
            } else if (pre) {  // on_infected, :41-50
                if (hour - t0v >= P.pre_symptomatic_duration) sev = bernoulli(a, P.thr_severe) ? SEV_SEVERE : SEV_MILD;
            }
        }
    }
    // re-pack
    s = (s & ~(ST_STATE_MASK | (3u << ST_SEV_SHIFT) | (ST_DAY_MAX << ST_DAY_SHIFT))) | state | (sev << ST_SEV_SHIFT) | (day << ST_DAY_SHIFT);
    uint32_t prop = 0;
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
        if (cell_byte(P, s) != cell_byte(P, s0)) prop |= PROP_DIRTY;
    }
    if (tx != x || ty != y) {
        prop |= PROP_MOVE | ((uint32_t)ty << CELL_BITS) | (uint32_t)tx;
        const uint32_t stamp = hour - D.clock->epoch_base + 1u;
        const uint32_t id_mask = (1u << P.id_bits) - 1u;
        atomicMax(&D.claim[P.cell_offset(tx, ty)], (stamp << P.id_bits) | (id_mask - i));
    }
    if (prop) prop |= (cell_byte(P, s) - 1u) << PROP_BYTE_SHIFT;
    st_stream(D.prop + i, prop);
}

// k_commit: four consecutive agents per thread.  The proposal and cell words arrive as two 128-bit loads, the (up to) four
// claim words are requested together before anything is stored, so a warp has 4x the memory requests in flight of a
// one-agent-per-thread version (the kernel is latency-bound: ~25 instructions per agent, two dependent round trips).
#ifndef EPI_CPF
#define EPI_CPF 0u  // L2 prefetch distance in agents (0: none -- with four agents per thread it no longer pays: 97 us vs 99-101 us)
#endif
__device__ __forceinline__ uint4 ld_stream4(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(uint32_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__global__ void __launch_bounds__(256) k_commit(Params P, DevPtrs D, uint32_t hour_offset) {
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    const uint32_t hour = D.clock->hour_base + hour_offset;
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // the hour's Counts row = sum of the running totals' copies (all k_hour blocks of this hour have finished)
#pragma unroll
        for (uint32_t c = 0; c < 6; ++c) {
            uint32_t v = threadIdx.x < TOT_COPIES ? D.tot[threadIdx.x * 8u + c] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (threadIdx.x == 0) D.counts[(size_t)(hour - D.clock->ring_base) * 8 + c] = v;
        }
    }
    if (i0 >= P.n) return;
    if (EPI_CPF != 0u && (threadIdx.x & 1u) == 0 && i0 + EPI_CPF < P.n) {  // one prefetch per 32-byte sector
        prefetch_l2(D.prop + i0 + EPI_CPF);
        prefetch_l2(D.cell + i0 + EPI_CPF);
    }
    const bool full = i0 + 3u < P.n;  // P.n need not be a multiple of 4
    uint32_t prop[4], c0[4];
    if (full) {
        const uint4 p4 = ld_stream4(D.prop + i0), c4 = ld_stream4(D.cell + i0);
        prop[0] = p4.x; prop[1] = p4.y; prop[2] = p4.z; prop[3] = p4.w;
        c0[0] = c4.x; c0[1] = c4.y; c0[2] = c4.z; c0[3] = c4.w;
    } else {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const bool in = i0 + j < P.n;
            prop[j] = in ? D.prop[i0 + j] : 0u;
            c0[j] = in ? D.cell[i0 + j] : 0u;
        }
    }
    if ((prop[0] | prop[1] | prop[2] | prop[3]) == 0) return;
    const uint32_t stamp = hour - D.clock->epoch_base + 1u;
    const uint32_t id_mask = (1u << P.id_bits) - 1u;
    // every claim word first: one round trip for the four agents
    uint32_t cl[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        cl[j] = 0;
        if (prop[j] & PROP_MOVE) cl[j] = __ldcg(D.claim + P.cell_offset(prop[j] & PROP_CELL_MASK));
    }
    bool moved = false;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        if (prop[j] == 0) continue;
        const uint32_t byte = (prop[j] >> PROP_BYTE_SHIFT) + 1u;
        const size_t at = P.cell_offset(c0[j]);
        if (prop[j] & PROP_MOVE) {
            const uint32_t tc = prop[j] & PROP_CELL_MASK;
            // lowest id among the claimants: upcoming.entry(new).or_insert (allocation_map.rs:93-98)
            if (cl[j] == ((stamp << P.id_bits) | (id_mask - (i0 + j)))) {
                D.grid[at] = 0;
                D.grid[P.cell_offset(tc)] = (uint8_t)byte;
                c0[j] = tc;
                moved = true;
                continue;
            }
            // lost: stays at old_cell (allocation_map.rs:99-102); still refresh the byte if it changed
        }
        if (prop[j] & PROP_DIRTY) D.grid[at] = (uint8_t)byte;
    }
    if (moved) {
        if (full) st_stream4(D.cell + i0, c0[0], c0[1], c0[2], c0[3]);
        else {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (i0 + j < P.n) D.cell[i0 + j] = c0[j];
        }
    }
}

// Four agents per thread (one 128-bit load); the six Counts columns travel as 10-bit fields of one 64-bit word through the warp
// reduction (a warp holds at most 128 agents), so the whole recount costs ten shuffles and six shared atomics per warp.
__global__ void __launch_bounds__(256) k_sleep(Params P, DevPtrs D, uint32_t hour_offset) {
    __shared__ uint32_t s_cnt[6];
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    const uint32_t hour = D.clock->hour_base + hour_offset;
    uint32_t st[4] = {ST_ABSENT, ST_ABSENT, ST_ABSENT, ST_ABSENT};
    const bool full = i0 + 3u < P.n;
    if (full) {
        const uint4 v = *reinterpret_cast<const uint4*>(D.st + i0);
        st[0] = v.x; st[1] = v.y; st[2] = v.z; st[3] = v.w;
    } else {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (i0 + j < P.n) st[j] = D.st[i0 + j];
    }
    unsigned long long packed = 0;
    bool changed = false;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        uint32_t s = st[j];
        if ((s & ST_STATE_MASK) != ST_ABSENT && ((s >> ST_WS_SHIFT) & 3u) != WS_STAFF && (s & ST_AREA_MASK) != (AK_HOME << ST_AREA_SHIFT)) {
            s = (s & ~ST_AREA_MASK) | (AK_HOME << ST_AREA_SHIFT);  // citizen/mod.rs:244-248
            st[j] = s;
            changed = true;
        }
        const uint32_t cat = count_category(s);
        if (cat < 6u) packed += 1ull << (10u * cat);
    }
    if (changed) {
        if (full) *reinterpret_cast<uint4*>(D.st + i0) = make_uint4(st[0], st[1], st[2], st[3]);
        else {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (i0 + j < P.n) D.st[i0 + j] = st[j];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xFFFFFFFFu, packed, o);
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (uint32_t c = 0; c < 6; ++c) {
            const uint32_t v = (uint32_t)(packed >> (10u * c)) & 1023u;
            if (v) atomicAdd(&s_cnt[c], v);
        }
    }
    __syncthreads();
    uint32_t* out_row = D.counts + (size_t)(hour - D.clock->ring_base) * 8;
    if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&out_row[threadIdx.x], s_cnt[threadIdx.x]);
}

__global__ void k_set_clock(Clock* clock, Clock value) { *clock = value; }

// absolute recount into copy 0 of the running totals (the other copies must have been zeroed): create / reset / set_state
__global__ void __launch_bounds__(256) k_recount(Params P, const uint32_t* __restrict__ st, uint32_t* __restrict__ tot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    block_count(i < P.n ? count_category(st[i]) : 6u, tot);
}

__global__ void __launch_bounds__(256) k_lock(Params P, uint32_t* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t s = st[i];
    if ((s & ST_STATE_MASK) != ST_ABSENT && ((s >> ST_WS_SHIFT) & 3u) != WS_ESSENTIAL && !(s & ST_ISO)) st[i] = s | ST_ISO;
}
__global__ void __launch_bounds__(256) k_unlock(Params P, uint32_t* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t s = st[i];
    if ((s & ST_STATE_MASK) != ST_ABSENT && (s & ST_ISO)) st[i] = s & ~ST_ISO;
}
__global__ void __launch_bounds__(256) k_vaccinate(Params P, uint32_t* __restrict__ st, uint64_t thr, uint32_t hour) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t s = st[i];
    if ((s & ST_STATE_MASK) == ST_S && !(s & ST_VACC) && bernoulli(philox_draw(P.seed, i, hour, DOM_VACCINATE, 0), thr)) st[i] = s | ST_VACC;
}

// rebuild the occupancy grid from agent state (after epi_set_state / epi_reset)
__global__ void __launch_bounds__(256) k_build_grid(Params P, const uint32_t* __restrict__ cell, const uint32_t* __restrict__ st, uint8_t* __restrict__ grid,
                                                     uint32_t* __restrict__ collisions) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    if ((st[i] & ST_STATE_MASK) == ST_ABSENT) return;
    const uint32_t c = cell[i];
    const size_t at = P.cell_offset(c);
    // byte-wide check-and-set through the containing word
    uint32_t* word = (uint32_t*)(grid + (at & ~(size_t)3));
    const uint32_t shift = (uint32_t)(at & 3) * 8u;
    const uint32_t old = atomicOr(word, cell_byte(P, st[i]) << shift);
    if ((old >> shift) & 0xFFu) atomicAdd(collisions, 1u);
}

// epi_set_state: the caller's arrays were copied to the device as they are (cell_x / cell_y into scratch, house / office INDICES
// into home[] / work[]); pack the cell, turn the indices into origins, validate.  bad[0] counts out-of-range entries.
__global__ void __launch_bounds__(256) k_import_state(Params P, DevPtrs D, const int32_t* __restrict__ cx, const int32_t* __restrict__ cy, uint32_t n_houses,
                                                       uint32_t n_offices, uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t s = D.st[i];
    const uint32_t ws = (s >> ST_WS_SHIFT) & 3u;
    const int x = cx[i], y = cy[i];
    const uint32_t h = D.home[i], w = D.work[i];
    if (x < 0 || y < 0 || (uint32_t)x >= P.pitch || (uint32_t)y >= P.rows || h >= n_houses || (ws != WS_NA && w >= n_offices)) {
        atomicAdd(bad, 1u);
        return;
    }
    D.cell[i] = ((uint32_t)y << CELL_BITS) | (uint32_t)x;
    D.home[i] = ((uint32_t)(P.housing().sy + 2 * (int)(h / (uint32_t)P.house_nx)) << CELL_BITS) | (uint32_t)(P.housing().sx + 2 * (int)(h % (uint32_t)P.house_nx));
    D.work[i] = ws == WS_NA ? 0u : ((uint32_t)(P.work.sy + 10 * (int)(w / (uint32_t)P.office_nx)) << CELL_BITS) | (uint32_t)(P.work.sx + 10 * (int)(w % (uint32_t)P.office_nx));
    D.reg[i] = (uint32_t)P.region | ((uint32_t)P.region << 8);
    D.prop[i] = 0;
}

// ---- launchers ---------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint32_t n) { return (n + 255u) / 256u; }
void launch_import_state(const Params& P, const DevPtrs& D, const int32_t* cx, const int32_t* cy, uint32_t n_houses, uint32_t n_offices, uint32_t* bad, cudaStream_t s) {
    k_import_state<<<blocks_for(P.n), 256, 0, s>>>(P, D, cx, cy, n_houses, n_offices, bad);
}

void launch_hospital_scan(const Params& P, const DevPtrs& D, cudaStream_t s) {
    const Rect h = P.hospital();
    const uint64_t total = (uint64_t)(h.ex - h.sx + 1) * (uint64_t)(h.ey - h.sy + 1);
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    if (blocks == 0) blocks = 1;
    k_hospital_scan<<<blocks, 256, 0, s>>>(P, D.grid, D.hosp_first);
}
template <bool INJECT>
static void launch_hour_t(const Params& P, const DevPtrs& D, uint32_t hour_of_day, uint32_t hour_offset, cudaStream_t s) {
    const unsigned b = (P.n + EPI_HBS - 1u) / EPI_HBS;
    switch (hour_of_day) {
        case 0: k_hour<KIND_START, INJECT><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;
        case 23: k_hour<KIND_END, INJECT><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;
        case 7: k_hour<KIND_MOVE, INJECT, 7><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;
        case 8: k_hour<KIND_MOVE, INJECT, 8><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;
        case 12: k_hour<KIND_MOVE, INJECT, 12><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;
        case 16: k_hour<KIND_MOVE, INJECT, 16><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;
        case 17: k_hour<KIND_MOVE, INJECT, 17><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;
        default: k_hour<KIND_MOVE, INJECT, 9><<<b, EPI_HBS, 0, s>>>(P, D, hour_offset); break;  // 9..11, 13..15, 18..22 (1..6 are k_sleep's)
    }
}
void launch_hour(const Params& P, const DevPtrs& D, uint32_t hour_of_day, uint32_t hour_offset, bool inject, cudaStream_t s) {
    if (inject) launch_hour_t<true>(P, D, hour_of_day, hour_offset, s);
    else launch_hour_t<false>(P, D, hour_of_day, hour_offset, s);
}
void launch_recount(const Params& P, const DevPtrs& D, cudaStream_t s) { k_recount<<<blocks_for(P.n), 256, 0, s>>>(P, D.st, D.tot); }
void launch_set_clock(Clock* clock, const Clock& value, cudaStream_t s) { k_set_clock<<<1, 1, 0, s>>>(clock, value); }
void launch_commit(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s) { k_commit<<<blocks_for((P.n + 3u) / 4u), 256, 0, s>>>(P, D, hour_offset); }
void launch_sleep(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s) { k_sleep<<<blocks_for((P.n + 3u) / 4u), 256, 0, s>>>(P, D, hour_offset); }
void launch_lock(const Params& P, const DevPtrs& D, cudaStream_t s) { k_lock<<<blocks_for(P.n), 256, 0, s>>>(P, D.st); }
void launch_unlock(const Params& P, const DevPtrs& D, cudaStream_t s) { k_unlock<<<blocks_for(P.n), 256, 0, s>>>(P, D.st); }
void launch_vaccinate(const Params& P, const DevPtrs& D, uint64_t thr, uint32_t hour, cudaStream_t s) { k_vaccinate<<<blocks_for(P.n), 256, 0, s>>>(P, D.st, thr, hour); }
void launch_build_grid(const Params& P, const DevPtrs& D, uint32_t* collisions, cudaStream_t s) {
    k_build_grid<<<blocks_for(P.n), 256, 0, s>>>(P, D.cell, D.st, D.grid, collisions);
}

}  // namespace epi
