// sm_100a kernels of the per-hour agent step.  HBM/L2-bound integer work: no tensor cores.
//
//   k_hospital_scan   first vacant hospital cell in row-major order        (allocation_map.rs:144-147)
//   k_hour<INJECT>    one agent per thread: routine, movement proposal against the start-of-hour grid, disease
//                     transition, Counts; atomicMax claim on the target cell (citizen/mod.rs:227-432,
//                     default_disease_handler.rs:31-103, counts.rs:126-140)
//   k_commit          lowest-id claimant moves, loser stays; grid bytes updated in place (allocation_map.rs:93-102,131-134)
//   k_sleep           hours 1..6: current_area := home (citizen/mod.rs:244-248) + Counts
//   k_lock / k_unlock / k_vaccinate   intervention sweeps (allocation_map.rs:349-387)
//
// Synchronous-update argument (why the grid can be updated in place): every proposal targets a cell that was vacant
// at the start of the hour (goto_area / move_agent_from / goto_hospital / deceased all go through
// CitizenLocationMap::move_agent or an is_cell_vacant filter), every cell that is cleared was occupied at the start
// of the hour, so the set of written-to-occupied and written-to-vacant cells are disjoint, and all reads of the grid
// happen in k_hour, all writes in k_commit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "layout.h"
#include "philox.cuh"

namespace epi {

// Moore neighbourhood in the reference's iterator order (geography/point.rs:59)
__constant__ int c_dx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
__constant__ int c_dy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};

__device__ __forceinline__ bool rect_contains(const Rect& r, int x, int y) { return r.sx <= x && r.ex >= x && r.sy <= y && r.ey >= y; }
__device__ __forceinline__ bool rect_eq(const Rect& a, const Rect& b) { return a.sx == b.sx && a.sy == b.sy && a.ex == b.ex && a.ey == b.ey; }

__device__ __forceinline__ Rect house_rect(const Params& P, uint32_t idx) {
    const int hx = (int)(idx % (uint32_t)P.house_nx), hy = (int)(idx / (uint32_t)P.house_nx);
    Rect r;
    r.sx = P.housing.sx + 2 * hx; r.sy = P.housing.sy + 2 * hy; r.ex = r.sx + 1; r.ey = r.sy + 1;
    return r;
}
__device__ __forceinline__ Rect office_rect(const Params& P, uint32_t idx) {
    const int ox = (int)(idx % (uint32_t)P.office_nx), oy = (int)(idx / (uint32_t)P.office_nx);
    Rect r;
    r.sx = P.work.sx + 10 * ox; r.sy = P.work.sy + 10 * oy; r.ex = r.sx + 9; r.ey = r.sy + 9;
    return r;
}

// Disease::get_current_transmission_rate as a class (common/src/disease/mod.rs:88-95); d = (day + immunity) as u32, wrapping
__device__ __forceinline__ uint32_t rate_class(const Params& P, uint32_t d) {
    if (P.regular_start < d && d <= P.high_start) return 1;
    if (P.high_start < d && d <= P.last_day) return 2;
    return 0;
}
// what other agents can see of this agent: occupied + Citizen::get_infection_transmission_rate for infected && !hospitalized
__device__ __forceinline__ uint32_t cell_byte(const Params& P, uint32_t s) {
    if ((s & ST_STATE_MASK) == ST_I && !(s & ST_HOSP)) {
        const int day = (int)(s >> ST_DAY_SHIFT), imm = (int)((s >> ST_IMM_SHIFT) & 7u) - 2;
        return 1u + rate_class(P, (uint32_t)(day + imm));
    }
    return 1u;
}

template <bool INJECT>
struct Draws {
    uint64_t seed;
    uint32_t agent, hour;
    const uint64_t* row;
    __device__ __forceinline__ uint64_t get(uint32_t slot) const {
        if (INJECT) return row[slot];
        return philox_draw(seed, agent, hour, DOM_STEP, slot);
    }
    // slots 2k and 2k+1 with one Philox call
    __device__ __forceinline__ void pair(uint32_t even_slot, uint64_t& a, uint64_t& b) const {
        if (INJECT) { a = row[even_slot]; b = row[even_slot + 1]; return; }
        const U4 o = philox4x32_10(agent, hour, even_slot >> 1, DOM_STEP, (uint32_t)seed, (uint32_t)(seed >> 32));
        a = (uint64_t)o.x | ((uint64_t)o.y << 32);
        b = (uint64_t)o.z | ((uint64_t)o.w << 32);
    }
};
enum : uint32_t { SLOT_PX = 0, SLOT_PY = 1, SLOT_PICK = 2, SLOT_A = 3, SLOT_B = 4, SLOT_EXPOSE0 = 8 };

struct Mover {
    const Params& P;
    const uint8_t* __restrict__ grid;
    __device__ __forceinline__ bool vacant(int x, int y) const { return grid[(size_t)y * P.pitch + (size_t)x] == 0; }
    __device__ __forceinline__ bool in_grid(int x, int y) const { return x >= 0 && y >= 0 && x < P.grid_size && y < P.grid_size; }
};

// Area::get_random_point (geography/area.rs:76-81)
template <bool INJECT>
__device__ __forceinline__ void random_point(const Draws<INJECT>& dr, const Rect& r, int& px, int& py) {
    uint64_t a, b;
    dr.pair(SLOT_PX, a, b);
    px = r.sx + (int)mulhi64(a, (uint64_t)(r.ex - r.sx + 1));
    py = r.sy + (int)mulhi64(b, (uint64_t)(r.ey - r.sy + 1));
}

// Citizen::move_agent_from (citizen/mod.rs:415-432).  pick_draw: slot SLOT_PICK
template <bool INJECT>
__device__ __forceinline__ void walk(const Mover& mv, const Draws<INJECT>& dr, const Rect& area, bool can_move, int x, int y, int& tx, int& ty) {
    tx = x; ty = y;
    if (!can_move) return;
    int lx = x, ly = y;
    if (!rect_contains(area, x, y)) random_point(dr, area, lx, ly);
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int nx = lx + c_dx[j], ny = ly + c_dy[j];
        if (rect_contains(area, nx, ny) && mv.in_grid(nx, ny) && mv.vacant(nx, ny)) mask |= 1u << j;
    }
    const int k = __popc(mask);
    if (k == 0) return;
    const uint32_t idx = (uint32_t)mulhi64(dr.get(SLOT_PICK), (uint64_t)k);
    const int j = (int)__fns(mask, 0, (int)idx + 1);
    tx = lx + c_dx[j]; ty = ly + c_dy[j];
}

__device__ __forceinline__ void block_count(uint32_t cat, uint32_t* __restrict__ out_row) {
    // warp-shuffle/ballot reduction -> shared -> one atomic per category per block (counts.rs:126-140)
    __shared__ uint32_t s_cnt[6];
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (uint32_t c = 0; c < 6; ++c) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, cat == c);
        if (lane == 0 && b) atomicAdd(&s_cnt[c], (uint32_t)__popc(b));
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&out_row[threadIdx.x], s_cnt[threadIdx.x]);
}
__device__ __forceinline__ uint32_t count_category(uint32_t s) {
    const uint32_t st = s & ST_STATE_MASK;  // order of the CSV columns: S,E,I,H,R,D
    return st == ST_S ? 0u : st == ST_E ? 1u : st == ST_I ? ((s & ST_HOSP) ? 3u : 2u) : st == ST_R ? 4u : 5u;
}

__global__ void __launch_bounds__(256) k_hospital_scan(Params P, const uint8_t* __restrict__ grid, uint32_t* __restrict__ hosp_first) {
    const Rect h = P.hospital[P.hospital_gen];
    const uint32_t w = (uint32_t)(h.ex - h.sx + 1), nh = (uint32_t)(h.ey - h.sy + 1);
    const uint64_t total = (uint64_t)w * nh;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (uint64_t)gridDim.x * blockDim.x) {
        if (r >= *(volatile uint32_t*)hosp_first) return;  // ranks only grow along the stride: nothing better ahead
        const uint32_t x = (uint32_t)h.sx + (uint32_t)(r % w), y = (uint32_t)h.sy + (uint32_t)(r / w);
        if (grid[(size_t)y * P.pitch + x] == 0) { atomicMin(hosp_first, (uint32_t)r); return; }
    }
}

template <bool INJECT>
__global__ void __launch_bounds__(256) k_hour(Params P, DevPtrs D, uint32_t hour_offset) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t hour = D.clock->hour_base + hour_offset;
    const uint32_t h = hour % 24u;
    uint32_t cat = 6;
    if (i < P.n) {
        const uint32_t s0 = D.st[i];
        const uint32_t c0 = D.cell[i];
        const int x = (int)(c0 & CELL_XMASK), y = (int)(c0 >> CELL_BITS);
        uint32_t s = s0;
        int tx = x, ty = y;
        const Mover mv{P, D.grid};
        Draws<INJECT> dr{P.seed, i, hour, INJECT ? D.draws + (size_t)i * 16 : nullptr};
        const uint32_t ws = (s >> ST_WS_SHIFT) & 3u;
        uint32_t state = s & ST_STATE_MASK, sev = (s >> ST_SEV_SHIFT) & 3u, day = s >> ST_DAY_SHIFT;
        const int imm = (int)((s >> ST_IMM_SHIFT) & 7u) - 2;
        const Rect home = house_rect(P, D.home[i] & INDEX_MASK);

        if (h == 0) {
            // ROUTINE_START_TIME: increment_infection_day + hospitalize (citizen/mod.rs:240-243, :351-365)
            if (state == ST_I) day = min(day + 1u, ST_DAY_MAX);
            if (!(s & ST_HOSP) && state == ST_I && sev == SEV_SEVERE && ((P.hospitalize_mask >> rate_class(P, (uint32_t)((int)day + imm))) & 1u)) {
                const uint32_t first = *D.hosp_first;
                if (first != HOSP_NONE) {  // goto_hospital: every admitted agent targets the same first vacant cell
                    const Rect hr = P.hospital[P.hospital_gen];
                    const uint32_t w = (uint32_t)(hr.ex - hr.sx + 1);
                    tx = hr.sx + (int)(first % w); ty = hr.sy + (int)(first / w);
                    s |= ST_HOSP;
                } else {  // hospital full: try a random point of the own house
                    int px, py;
                    random_point(dr, home, px, py);
                    if (mv.vacant(px, py)) { tx = px; ty = py; }
                }
            }
        } else if (h == 23) {
            // ROUTINE_END_TIME: Citizen::deceased + on_routine_end (citizen/mod.rs:397-413, default_disease_handler.rs:88-103)
            if (state == ST_I) {
                if ((sev == SEV_ASYM && day == 9u) || (sev == SEV_MILD && day == 12u)) state = ST_R;
                else if (sev == SEV_SEVERE && day == P.last_day) state = bernoulli(dr.get(SLOT_A), P.thr_death) ? ST_D : ST_R;
            }
            if (state == ST_R) {  // every recovered agent, every day
                int px, py;
                random_point(dr, home, px, py);
                if (mv.vacant(px, py)) { tx = px; ty = py; }
            }
            if (state == ST_R || state == ST_D) { s &= ~ST_HOSP; sev = 0; day = 0; }
        } else {
            // perform_movements (citizen/mod.rs:257-349); h in 7..22 here (sleep hours use k_sleep)
            const uint32_t kind0 = (s >> ST_AREA_SHIFT) & 7u;
            const bool symptomatic = state == ST_I && (sev == SEV_MILD || sev == SEV_SEVERE);
            const bool can_move = !(symptomatic || (s & ST_HOSP) || state == ST_D || (s & ST_ISO));  // citizen/mod.rs:452-454
            const Rect workr = ws == WS_NA ? home : office_rect(P, D.work[i] & INDEX_MASK);
            auto rect_of = [&](uint32_t kind) -> Rect {
                switch (kind) {
                    case AK_HOME: return home;
                    case AK_WORK: return workr;
                    case AK_TRANSPORT: return P.transport;
                    case AK_HOUSING: return P.housing;
                    case AK_HOSPITAL0: return P.hospital[0];
                    default: return P.hospital[1];
                }
            };
            const Rect cur0 = rect_of(kind0);
            uint32_t kind = kind0;
            bool dynamics = true;
            // Citizen::goto_area (citizen/mod.rs:367-395)
            auto goto_area = [&](const Rect& target) {
                bool override_movement = false;
                if (ws == WS_NORMAL || ws == WS_ESSENTIAL)
                    override_movement = rect_contains(workr, x, y) && rect_eq(target, home) && symptomatic;
                if (!can_move && !override_movement) return;
                if (ws != WS_NA) {
                    int px, py;
                    random_point(dr, target, px, py);
                    if (mv.vacant(px, py)) { tx = px; ty = py; }
                } else {
                    walk(mv, dr, cur0, can_move, x, y, tx, ty);
                }
            };
            if (ws == WS_NORMAL || ws == WS_ESSENTIAL) {
                if (h == 7 || h == 17) {
                    if (s & ST_PT) { goto_area(P.transport); kind = AK_TRANSPORT; }
                    else walk(mv, dr, cur0, can_move, x, y, tx, ty);
                } else if (h == 8) { goto_area(workr); kind = AK_WORK; }
                else if (h == 16) { goto_area(home); kind = AK_HOME; }
                else walk(mv, dr, cur0, can_move, x, y, tx, ty);
            } else if (ws == WS_NA) {
                if (h == 8) { goto_area(P.housing); kind = AK_HOUSING; }
                else if (h == 12) { goto_area(home); kind = AK_HOME; }
                else walk(mv, dr, cur0, can_move, x, y, tx, ty);
            } else {  // HospitalStaff { work_start_at }
                uint32_t wsa = D.wsa[i];
                const uint32_t since = hour >= wsa ? hour - wsa : 0u;  // saturating_sub
                if (since == 24u * 14u) { s |= ST_WQ; dynamics = false; }
                else if (since == 24u * 14u * 2u) {
                    goto_area(home); kind = AK_HOME;
                    D.wsa[i] = hour + 24u * 14u;
                    dynamics = false;
                } else if (h == 8) {
                    const Rect hr = P.hospital[P.hospital_gen];
                    if (!rect_eq(cur0, hr) && wsa <= hour) {
                        goto_area(hr); kind = P.hospital_gen ? AK_HOSPITAL1 : AK_HOSPITAL0;
                        D.wsa[i] = hour;
                    }
                    s &= ~ST_WQ;
                } else if (h == 16) { s |= ST_WQ; }
                else if (!(s & ST_WQ) && can_move) walk(mv, dr, cur0, can_move, x, y, tx, ty);
            }
            s = (s & ~ST_AREA_MASK) | (kind << ST_AREA_SHIFT);
            if (dynamics) {
                // DiseaseStateMachine::next at the proposed cell (disease_state_machine.rs:53-70)
                if (state == ST_S) {
                    if (!(s & (ST_WQ | ST_VACC))) {  // on_susceptible, default_disease_handler.rs:64-86
                        const Rect cur = kind == kind0 ? cur0 : rect_of(kind);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int nx = tx + c_dx[j], ny = ty + c_dy[j];
                            if (!(rect_contains(cur, nx, ny) && mv.in_grid(nx, ny))) continue;
                            const uint32_t b = D.grid[(size_t)ny * P.pitch + (size_t)nx];
                            if (b >= 2u && bernoulli(dr.get(SLOT_EXPOSE0 + j), P.thr_rate[b - 1u])) {
                                state = ST_E;
                                D.t0[i] = hour;
                                break;
                            }
                        }
                    }
                } else if (state == ST_E) {  // on_exposed, :52-62
                    uint64_t da, db_unused;
                    dr.pair(SLOT_PICK, db_unused, da);  // slot 3 = high pair of block 1
                    const int f = (int)mulhi64(da, 3ull) - 1;
                    if (hour - D.t0[i] >= (uint32_t)((int)P.exposed_duration + f)) {
                        const bool symptoms = bernoulli(dr.get(SLOT_B), P.thr_symptomatic);
                        state = ST_I; day = 0;
                        sev = symptoms ? SEV_PRE : SEV_ASYM;
                        if (symptoms) D.t0[i] = hour;
                    }
                } else if (state == ST_I) {  // on_infected, :41-50
                    if (sev == SEV_PRE && hour - D.t0[i] >= P.pre_symptomatic_duration)
                        sev = bernoulli(dr.get(SLOT_A), P.thr_severe) ? SEV_SEVERE : SEV_MILD;
                }
            }
        }
        // re-pack
        s = (s & ~(ST_STATE_MASK | (3u << ST_SEV_SHIFT) | (ST_DAY_MAX << ST_DAY_SHIFT))) | state | (sev << ST_SEV_SHIFT) | (day << ST_DAY_SHIFT);
        if (s != s0) D.st[i] = s;
        const uint32_t b_old = cell_byte(P, s0), b_new = cell_byte(P, s);
        uint32_t prop = 0;
        if (b_new != b_old) prop |= PROP_DIRTY;
        if (tx != x || ty != y) {
            prop |= PROP_MOVE | ((uint32_t)ty << CELL_BITS) | (uint32_t)tx;
            const uint32_t stamp = hour - D.clock->epoch_base + 1u;
            const uint32_t id_mask = (1u << P.id_bits) - 1u;
            atomicMax(&D.claim[(size_t)ty * P.pitch + (size_t)tx], (stamp << P.id_bits) | (id_mask - i));
        }
        if (prop) prop |= (b_new - 1u) << PROP_BYTE_SHIFT;
        D.prop[i] = prop;
        cat = count_category(s);
    }
    block_count(cat, D.counts + (size_t)(hour - D.clock->ring_base) * 8);
}

__global__ void __launch_bounds__(256) k_commit(Params P, DevPtrs D, uint32_t hour_offset) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t prop = D.prop[i];
    if (prop == 0) return;
    const uint32_t byte = (prop >> PROP_BYTE_SHIFT) + 1u;
    const uint32_t c0 = D.cell[i];
    size_t at = (size_t)(c0 >> CELL_BITS) * P.pitch + (c0 & CELL_XMASK);
    if (prop & PROP_MOVE) {
        const uint32_t hour = D.clock->hour_base + hour_offset;
        const uint32_t stamp = hour - D.clock->epoch_base + 1u;
        const uint32_t id_mask = (1u << P.id_bits) - 1u;
        const uint32_t tc = prop & PROP_CELL_MASK;
        const size_t tat = (size_t)(tc >> CELL_BITS) * P.pitch + (tc & CELL_XMASK);
        if (D.claim[tat] == ((stamp << P.id_bits) | (id_mask - i))) {  // lowest id among claimants: upcoming.entry(new).or_insert
            D.grid[at] = 0;
            D.grid[tat] = (uint8_t)byte;
            D.cell[i] = tc;
            return;
        }
        // lost: stays at old_cell (allocation_map.rs:99-102); still refresh the byte if it changed
    }
    if (prop & PROP_DIRTY) D.grid[at] = (uint8_t)byte;
}

__global__ void __launch_bounds__(256) k_sleep(Params P, DevPtrs D, uint32_t hour_offset) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t hour = D.clock->hour_base + hour_offset;
    uint32_t cat = 6;
    if (i < P.n) {
        uint32_t s = D.st[i];
        if (((s >> ST_WS_SHIFT) & 3u) != WS_STAFF && (s & ST_AREA_MASK) != (AK_HOME << ST_AREA_SHIFT)) {
            s = (s & ~ST_AREA_MASK) | (AK_HOME << ST_AREA_SHIFT);
            D.st[i] = s;
        }
        cat = count_category(s);
    }
    block_count(cat, D.counts + (size_t)(hour - D.clock->ring_base) * 8);
}

__global__ void k_set_clock(Clock* clock, Clock value) { *clock = value; }

__global__ void __launch_bounds__(256) k_lock(Params P, uint32_t* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t s = st[i];
    if (((s >> ST_WS_SHIFT) & 3u) != WS_ESSENTIAL && !(s & ST_ISO)) st[i] = s | ST_ISO;
}
__global__ void __launch_bounds__(256) k_unlock(Params P, uint32_t* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t s = st[i];
    if (s & ST_ISO) st[i] = s & ~ST_ISO;
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
    const uint32_t c = cell[i];
    const size_t at = (size_t)(c >> CELL_BITS) * P.pitch + (c & CELL_XMASK);
    // byte-wide check-and-set through the containing word
    uint32_t* word = (uint32_t*)(grid + (at & ~(size_t)3));
    const uint32_t shift = (uint32_t)(at & 3) * 8u;
    const uint32_t old = atomicOr(word, cell_byte(P, st[i]) << shift);
    if ((old >> shift) & 0xFFu) atomicAdd(collisions, 1u);
}

// ---- launchers ---------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint32_t n) { return (n + 255u) / 256u; }

void launch_hospital_scan(const Params& P, const DevPtrs& D, cudaStream_t s) {
    const Rect h = P.hospital[P.hospital_gen];
    const uint64_t total = (uint64_t)(h.ex - h.sx + 1) * (uint64_t)(h.ey - h.sy + 1);
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    if (blocks == 0) blocks = 1;
    k_hospital_scan<<<blocks, 256, 0, s>>>(P, D.grid, D.hosp_first);
}
void launch_hour(const Params& P, const DevPtrs& D, uint32_t hour_offset, bool inject, cudaStream_t s) {
    if (inject) k_hour<true><<<blocks_for(P.n), 256, 0, s>>>(P, D, hour_offset);
    else k_hour<false><<<blocks_for(P.n), 256, 0, s>>>(P, D, hour_offset);
}
void launch_set_clock(Clock* clock, const Clock& value, cudaStream_t s) { k_set_clock<<<1, 1, 0, s>>>(clock, value); }
void launch_commit(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s) { k_commit<<<blocks_for(P.n), 256, 0, s>>>(P, D, hour_offset); }
void launch_sleep(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s) { k_sleep<<<blocks_for(P.n), 256, 0, s>>>(P, D, hour_offset); }
void launch_lock(const Params& P, const DevPtrs& D, cudaStream_t s) { k_lock<<<blocks_for(P.n), 256, 0, s>>>(P, D.st); }
void launch_unlock(const Params& P, const DevPtrs& D, cudaStream_t s) { k_unlock<<<blocks_for(P.n), 256, 0, s>>>(P, D.st); }
void launch_vaccinate(const Params& P, const DevPtrs& D, uint64_t thr, uint32_t hour, cudaStream_t s) { k_vaccinate<<<blocks_for(P.n), 256, 0, s>>>(P, D.st, thr, hour); }
void launch_build_grid(const Params& P, const DevPtrs& D, uint32_t* collisions, cudaStream_t s) {
    k_build_grid<<<blocks_for(P.n), 256, 0, s>>>(P, D.cell, D.st, D.grid, collisions);
}

}  // namespace epi
