// sm_100a kernels of the daily traveller exchange (north_star: "packs and unpacks migrating agents on device").
// Replaces the per-agent parts of CitizenLocationMap::simulate's traveller selection (allocation_map.rs:104-120),
// remove_migrators / remove_commuters (:165-212) and assimilate_migrators / assimilate_commuters (:214-277).
// The payload goes GPU -> GPU (NCCL all-to-allv over NVLink, driven by the caller); the host only sees counts and the
// small index lists it needs for the reference's sequential bookkeeping (occupancy heaps, free slots).
//
//   k_travel_flag_count / k_travel_scan / k_travel_scatter   ordered stream compaction of the leaving agents (ascending slot)
//   k_travel_pack        records -> send buffer (grouped by destination), agents removed from the region
//   k_travel_install     arrivals -> agent slots (Citizen::from_migrator / from_commuter, citizen/mod.rs:113-154)
//   k_travel_propose / k_travel_grant   placement rounds: distinct vacant cells of the arrival strip, lowest arrival index wins
#include <cuda_runtime.h>
#include <stdint.h>

#include "agent.cuh"
#include "kernels.h"
#include "layout.h"
#include "philox.cuh"

namespace epi {

// Citizen::can_move (citizen/mod.rs:452-454) on the packed state word
__device__ __forceinline__ bool can_move_word(uint32_t s) {
    const uint32_t state = s & ST_STATE_MASK, sev = (s >> ST_SEV_SHIFT) & 3u;
    const bool symptomatic = state == ST_I && sev >= SEV_MILD;
    return !(symptomatic || (s & (ST_HOSP | ST_ISO)) || state == ST_D);
}

// 0 = stays; otherwise destination region + 1 (commuters) or 1 (migrator candidate; the host allots destinations)
__device__ __forceinline__ uint32_t travel_flag(const Params& P, const TravelArgs& A, uint32_t i, uint32_t s, uint32_t reg) {
    if ((s & ST_STATE_MASK) == ST_ABSENT || !can_move_word(s)) return 0;
    const uint32_t home_reg = reg & 0xFFu, work_reg = (reg >> 8) & 0xFFu, self = (uint32_t)P.region;
    if (A.kind == TRAVEL_COMMUTE) {  // Citizen::is_commuter (citizen/mod.rs:488-495)
        if (A.hour_of_day == 7u) return work_reg != self ? work_reg + 1u : 0u;
        return home_reg != self ? home_reg + 1u : 0u;
    }
    // Citizen::can_migrate (citizen/mod.rs:456-466; the hour window is checked by the host) && gen_bool(percent_outgoing)
    if (home_reg != self || work_reg != self) return 0;
    return bernoulli(philox_draw(P.seed, i, A.hour, DOM_MIGRATE, 0), A.thr_outgoing) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_travel_flag_count(Params P, DevPtrs D, TravelArgs A, uint32_t* __restrict__ block_counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t f = i < P.n ? travel_flag(P, A, i, D.st[i], D.reg[i]) : 0u;
    const int n = __syncthreads_count(f != 0);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)n;
}

// exclusive scan of the block counts by ONE block (the lists are short; this is launched 3 times per simulated day)
__global__ void __launch_bounds__(1024) k_travel_scan(uint32_t* __restrict__ block_counts, uint32_t n_blocks, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < n_blocks ? block_counts[idx] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if ((int)lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (idx < n_blocks) block_counts[idx] = incl - v;  // exclusive prefix
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(256) k_travel_scatter(Params P, DevPtrs D, TravelArgs A, const uint32_t* __restrict__ block_offsets,
                                                         uint32_t* __restrict__ out_slots, uint32_t* __restrict__ out_dest) {
    __shared__ uint32_t warp_counts[8];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t f = i < P.n ? travel_flag(P, A, i, D.st[i], D.reg[i]) : 0u;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(0xFFFFFFFFu, f != 0);
    if (lane == 0) warp_counts[warp] = (uint32_t)__popc(b);
    __syncthreads();
    if (f) {
        uint32_t rank = __popc(b & ((1u << lane) - 1u));
        for (unsigned w = 0; w < warp; ++w) rank += warp_counts[w];
        const uint32_t at = block_offsets[blockIdx.x] + rank;
        out_slots[at] = i;
        out_dest[at] = f - 1u;
    }
}

// send_slots[j] (already grouped by destination by the host) -> record j; the agent leaves: cell vacated, slot emptied,
// Counts decremented (decrement_counts, allocation_map.rs:291-301)
__global__ void __launch_bounds__(256) k_travel_pack(Params P, DevPtrs D, const uint32_t* __restrict__ send_slots, uint32_t n_send, TravelRecord* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_send) return;
    const uint32_t i = send_slots[j];
    const uint32_t s = D.st[i];
    TravelRecord r;
    r.st = s; r.t0 = D.t0[i]; r.home = D.home[i]; r.work = D.work[i]; r.reg = D.reg[i]; r.slot = i; r.from = (uint32_t)P.region; r.pad = 0;
    out[j] = r;
    const uint32_t c = D.cell[i];
    D.grid[P.cell_offset(c)] = 0;
    D.st[i] = ST_ABSENT;
    D.prop[i] = 0;
    atomicSub(D.tot + count_category(s), 1u);
}

// arrival k -> slot in_slot[k].  Citizen::from_migrator / from_commuter (citizen/mod.rs:113-154): immunity, vaccinated,
// uses_public_transport and the disease state travel; hospitalized, isolated, work_quarantined reset; work status becomes
// NA (migrator) or Normal (commuter); current_area = the housing strip.  The cell is assigned by the placement rounds.
__global__ void __launch_bounds__(256) k_travel_install(Params P, DevPtrs D, TravelArgs A, const TravelRecord* __restrict__ in, uint32_t n_in,
                                                         const uint32_t* __restrict__ in_slot, const uint32_t* __restrict__ in_home,
                                                         const uint32_t* __restrict__ in_work) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_in) return;
    const TravelRecord r = in[k];
    const uint32_t i = in_slot[k];
    const uint32_t keep = ST_STATE_MASK | (3u << ST_SEV_SHIFT) | (7u << ST_IMM_SHIFT) | ST_VACC | ST_PT | (ST_DAY_MAX << ST_DAY_SHIFT);
    const uint32_t ws = A.kind == TRAVEL_MIGRATE ? WS_NA : WS_NORMAL;
    const uint32_t s = (r.st & keep) | (ws << ST_WS_SHIFT) | (AK_HOUSING << ST_AREA_SHIFT);
    const uint32_t self = (uint32_t)P.region;
    D.st[i] = s;
    D.t0[i] = r.t0;
    D.wsa[i] = 0;
    D.prop[i] = 0;
    if (A.kind == TRAVEL_MIGRATE) {
        D.home[i] = in_home[k];
        D.work[i] = in_work[k];
        D.reg[i] = self | (self << 8);
    } else {
        D.home[i] = r.home;
        const bool assign_office = A.hour == 7u;  // sic: the absolute hour (allocation_map.rs:260), i.e. the first day only
        D.work[i] = assign_office ? in_work[k] : r.work;
        D.reg[i] = (r.reg & 0xFFu) | ((assign_office ? self : ((r.reg >> 8) & 0xFFu)) << 8);
    }
    atomicAdd(D.tot + count_category(s), 1u);
}

// ---- placement rounds (select_starting_points, allocation_map.rs:339-347) ------------------------------------------
// Arrival k proposes in round a the cell drawn from Philox(seed, k, hour, DOM_ARRIVAL) block a, x in [sx, ex), y in [sy, ey).
__device__ __forceinline__ uint32_t arrival_cell(const Params& P, const TravelArgs& A, uint32_t k, uint32_t attempt) {
    const U4 o = philox4x32_10(k, A.hour, attempt, DOM_ARRIVAL, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    const Rect& r = A.kind == TRAVEL_MIGRATE ? P.housing() : P.transport();
    const uint32_t x = (uint32_t)r.sx + __umulhi(o.x, (uint32_t)(r.ex - r.sx));
    const uint32_t y = (uint32_t)r.sy + __umulhi(o.y, (uint32_t)(r.ey - r.sy));
    return (y << CELL_BITS) | x;
}
__device__ __forceinline__ uint32_t hash_cell(uint32_t c) { return (c * 0x9E3779B1u) >> 7; }

__global__ void __launch_bounds__(256) k_travel_propose(Params P, DevPtrs D, TravelArgs A, uint32_t n_in, uint32_t attempt, const uint8_t* __restrict__ placed,
                                                         uint32_t* __restrict__ table_keys, uint32_t* __restrict__ table_vals, uint32_t table_mask) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_in || placed[k]) return;
    const uint32_t c = arrival_cell(P, A, k, attempt);
    if (D.grid[P.cell_offset(c)] != 0) return;  // occupied (also by earlier rounds' winners)
    uint32_t slot = hash_cell(c) & table_mask;
    for (;;) {
        const uint32_t old = atomicCAS(&table_keys[slot], 0u, c + 1u);
        if (old == 0u || old == c + 1u) { atomicMin(&table_vals[slot], k); return; }
        slot = (slot + 1u) & table_mask;
    }
}
__global__ void __launch_bounds__(256) k_travel_grant(Params P, DevPtrs D, TravelArgs A, uint32_t n_in, uint32_t attempt, uint8_t* __restrict__ placed,
                                                       const uint32_t* __restrict__ in_slot, const uint32_t* __restrict__ table_keys,
                                                       const uint32_t* __restrict__ table_vals, uint32_t table_mask, uint32_t* __restrict__ pending) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_in || placed[k]) return;
    const uint32_t c = arrival_cell(P, A, k, attempt);
    uint32_t slot = hash_cell(c) & table_mask;
    bool won = false;
    for (;;) {
        const uint32_t key = table_keys[slot];
        if (key == 0u) break;
        if (key == c + 1u) { won = table_vals[slot] == k; break; }
        slot = (slot + 1u) & table_mask;
    }
    if (won) {
        const uint32_t i = in_slot[k];
        D.cell[i] = c;
        D.grid[P.cell_offset(c)] = (uint8_t)cell_byte(P, D.st[i]);
        placed[k] = 1;
    } else {
        atomicAdd(pending, 1u);
    }
}

// ---- launchers ---------------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint32_t n) { return (n + 255u) / 256u; }

void launch_travel_select(const Params& P, const DevPtrs& D, const TravelArgs& A, uint32_t* block_counts, uint32_t* total, uint32_t* out_slots,
                          uint32_t* out_dest, int phase, cudaStream_t s) {
    const unsigned nb = blocks_for(P.n);
    if (phase == 0) {
        k_travel_flag_count<<<nb, 256, 0, s>>>(P, D, A, block_counts);
        k_travel_scan<<<1, 1024, 0, s>>>(block_counts, nb, total);
    } else {
        k_travel_scatter<<<nb, 256, 0, s>>>(P, D, A, block_counts, out_slots, out_dest);
    }
}
void launch_travel_pack(const Params& P, const DevPtrs& D, const uint32_t* send_slots, uint32_t n_send, TravelRecord* out, cudaStream_t s) {
    if (n_send) k_travel_pack<<<blocks_for(n_send), 256, 0, s>>>(P, D, send_slots, n_send, out);
}
void launch_travel_install(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelRecord* in, uint32_t n_in, const uint32_t* in_slot,
                           const uint32_t* in_home, const uint32_t* in_work, cudaStream_t s) {
    if (n_in) k_travel_install<<<blocks_for(n_in), 256, 0, s>>>(P, D, A, in, n_in, in_slot, in_home, in_work);
}
void launch_travel_round(const Params& P, const DevPtrs& D, const TravelArgs& A, uint32_t n_in, uint32_t attempt, uint8_t* placed, const uint32_t* in_slot,
                         uint32_t* table_keys, uint32_t* table_vals, uint32_t table_mask, uint32_t* pending, cudaStream_t s) {
    cudaMemsetAsync(table_keys, 0, ((size_t)table_mask + 1) * sizeof(uint32_t), s);
    cudaMemsetAsync(table_vals, 0xFF, ((size_t)table_mask + 1) * sizeof(uint32_t), s);
    cudaMemsetAsync(pending, 0, sizeof(uint32_t), s);
    k_travel_propose<<<blocks_for(n_in), 256, 0, s>>>(P, D, A, n_in, attempt, placed, table_keys, table_vals, table_mask);
    k_travel_grant<<<blocks_for(n_in), 256, 0, s>>>(P, D, A, n_in, attempt, placed, in_slot, table_keys, table_vals, table_mask, pending);
}

}  // namespace epi
