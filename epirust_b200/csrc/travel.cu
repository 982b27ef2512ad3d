// sm_100a kernels of the daily traveller exchange (north_star: "packs and unpacks migrating agents on device").
// Replaces CitizenLocationMap::simulate's traveller selection (allocation_map.rs:104-120), remove_migrators /
// remove_commuters (:165-212), assimilate_migrators / assimilate_commuters (:214-277), the allotment of migrators to regions
// (engine_migration_plan.rs:51-77, migrators_by_engine.rs:34-55), commuters_by_region.rs:59-78 and the house / office
// occupancy heaps (grid.rs:47-80, 279-341).  Everything the reference does sequentially on the host is done here on the
// device with the same result; the host only learns the per-region record counts (it needs them for the collective).
//
// Leaving (epi_travel_pack), two launches:
//   k_travel_select    flag of every slot (warp ballots), counts per block of 1024 slots (and per destination for commuters);
//                      the last block to finish scans them and writes the plan: records per destination (migrators by position
//                      in slot order -- floor(share * total) from the front, the surplus stays; commuters by their work / home
//                      region) and the segment headers
//   k_travel_move_out  walks the ballots in slot order: record -> segment of the destination at its stable index; the agent
//                      leaves: cell vacated, slot pushed on the free stack, Counts decremented, its house / office lose an
//                      occupant (migrators); the last block moves the stack top and the population
// Arriving (epi_travel_unpack), two launches (five when houses / offices are assigned):
//   k_travel_gather    segments (one per source region) -> contiguous arrival list, ordered by source region (+ the rank of every
//                      arriving migrator among the working ones: they also take an office)
//   k_occ_*            "water filling": the pop sequence of the reference's BinaryHeap of areas for K arrivals, in parallel.
//                      The heap pops the least occupied area, ties to the greatest (start.x, start.y); a popped area
//                      returns with one more occupant.  So the pops run level by level: first every area of the lowest
//                      occupancy L in tie order, then every area whose occupancy was <= L + 1 in tie order, and so on.
//                      occ[] is stored in tie order, so "the q-th area with occupancy <= l" is a prefix count.
//   k_travel_place     cooperative: arrival k -> the k-th slot from the top of the free stack (Citizen::from_migrator /
//                      from_commuter), placement rounds until every arrival stands on a distinct vacant cell of the arrival strip
//                      (lowest arrival wins), stack / population bookkeeping, the Counts row of the exchange hour
// Nothing here needs the host: counts, errors (sticky in TravelVars.err) and the population live on the device; the host reads
// them when it collects the Counts rows.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "agent.cuh"
#include "kernels.h"
#include "layout.h"
#include "philox.cuh"

namespace cg = cooperative_groups;

namespace epi {

// Citizen::can_move (citizen/mod.rs:452-454) on the packed state word
__device__ __forceinline__ bool can_move_word(uint32_t s) {
    const uint32_t state = s & ST_STATE_MASK, sev = (s >> ST_SEV_SHIFT) & 3u;
    const bool symptomatic = state == ST_I && sev >= SEV_MILD;
    return !(symptomatic || (s & (ST_HOSP | ST_ISO)) || state == ST_D);
}

// 0 = stays; otherwise destination region + 1 (commuters) or 1 (migrator candidate; k_travel_plan allots destinations)
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

// tie order of the occupancy heaps: greatest (start.x, start.y) first (grid.rs:67-73)
__device__ __forceinline__ uint32_t house_rank(const Params& P, uint32_t origin) {
    const int hx = ((int)(origin & CELL_XMASK) - P.housing().sx) / 2, hy = ((int)((origin >> CELL_BITS) & CELL_XMASK) - P.housing().sy) / 2;
    return (uint32_t)((P.house_nx - 1 - hx) * P.house_ny + (P.house_ny - 1 - hy));
}
__device__ __forceinline__ uint32_t house_origin_of_rank(const Params& P, uint32_t rank) {
    const int hx = P.house_nx - 1 - (int)(rank / (uint32_t)P.house_ny), hy = P.house_ny - 1 - (int)(rank % (uint32_t)P.house_ny);
    return ((uint32_t)(P.housing().sy + 2 * hy) << CELL_BITS) | (uint32_t)(P.housing().sx + 2 * hx);
}
__device__ __forceinline__ uint32_t office_rank(const Params& P, uint32_t origin) {
    const int ox = ((int)(origin & CELL_XMASK) - P.work.sx) / 10, oy = ((int)((origin >> CELL_BITS) & CELL_XMASK) - P.work.sy) / 10;
    return (uint32_t)((P.office_nx - 1 - ox) * P.office_ny + (P.office_ny - 1 - oy));
}
__device__ __forceinline__ uint32_t office_origin_of_rank(const Params& P, uint32_t rank) {
    const int ox = P.office_nx - 1 - (int)(rank / (uint32_t)P.office_ny), oy = P.office_ny - 1 - (int)(rank % (uint32_t)P.office_ny);
    return ((uint32_t)(P.work.sy + 10 * oy) << CELL_BITS) | (uint32_t)(P.work.sx + 10 * ox);
}

// exclusive prefix of `v` over the block (blockDim.x <= 1024, multiple of 32); *block_total receives the sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t* block_total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if ((int)lane >= o) x += y;
    }
    __syncthreads();  // warp_sums may still be read by the previous call
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < n_warps ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if ((int)lane >= o) w += y;
        }
        warp_sums[lane] = w;  // inclusive prefix of the warp sums
    }
    __syncthreads();
    *block_total = warp_sums[n_warps - 1];
    return x - v + (warp ? warp_sums[warp - 1] : 0u);
}

// ---- leaving ------------------------------------------------------------------------------------------------------------
// ONE cooperative launch (k_travel_leave); grid-wide barriers separate its phases:
//   1  flag of every slot, evaluated once: warp ballots, a count per block of 1024 slots, commuters per destination
//   2  block 0: offsets of the blocks (scan of the counts) and the plan -- records per destination (migrators by position in slot
//      order: floor(share * total) from the front, the surplus stays; commuters by their work / home region), segment headers
//   3  the ballots again, in slot order: a migrator knows its destination and its index there from its position alone and leaves on
//      the spot; commuters are listed (slot, destination) in order
//   4  commuters: leavers per destination in every chunk of 256 list entries
//   5  commuters: stable index inside the destination's segment = leavers of that destination in earlier chunks + earlier in the
//      chunk; the agent leaves
//   6  the free-slot stack and the population move by the number of leavers
constexpr uint32_t SEL_BLOCK = 1024;  // slots per count / offset entry

// inclusive warp scan; returns the inclusive prefix of v
__device__ __forceinline__ uint32_t warp_inclusive(uint32_t v) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if ((int)lane >= o) v += y;
    }
    return v;
}
// the whole block: arr[0 .. n) -> its exclusive prefix, in place (every thread owns a contiguous chunk); returns the total
__device__ __forceinline__ uint32_t block_scan_array(uint32_t* arr, uint32_t n, uint32_t* warp_sums) {
    const uint32_t per = (n + blockDim.x - 1u) / blockDim.x;
    const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
    uint32_t sum = 0;
    for (uint32_t k = b; k < e; ++k) sum += arr[k];
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, warp_sums, &total);
    for (uint32_t k = b; k < e; ++k) { const uint32_t v = arr[k]; arr[k] = run; run += v; }
    return total;
}

// The level histograms of an occupancy heap -- areas per occupancy level 0 .. cap-1 in every block of 256 areas (bh) and overall (tot)
// -- follow every change of occ[] (an area at cap or more, or absent, is in no level), so an exchange never has to recount them.
// (tot: TOT_COPIES spread copies against same-address contention -- thousands of areas change level in one exchange; the level
// totals are the column sums)
__device__ __forceinline__ void heap_level_moved(uint32_t* __restrict__ bh, uint32_t* __restrict__ tot, uint32_t cap, uint32_t area, uint32_t from, uint32_t to) {
    const size_t row = (size_t)(area >> 8) * cap;
    uint32_t* t = tot + (size_t)(area & (TOT_COPIES - 1u)) * cap;
    if (from < cap) { atomicSub(&bh[row + from], 1u); atomicSub(&t[from], 1u); }
    if (to < cap) { atomicAdd(&bh[row + to], 1u); atomicAdd(&t[to], 1u); }
}
// (re)count them: reset / creation
__global__ void __launch_bounds__(256) k_occ_count_levels(const uint32_t* __restrict__ occ, uint32_t n, uint32_t cap, uint32_t* __restrict__ bh, uint32_t* __restrict__ tot) {
    __shared__ uint32_t h[OFFICE_CAP];
    for (uint32_t l = threadIdx.x; l < cap; l += blockDim.x) h[l] = 0;
    __syncthreads();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t o = r < n ? occ[r] : OCC_ABSENT;
    if (o < cap) atomicAdd(&h[o], 1u);
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < cap; l += blockDim.x) {
        bh[(size_t)blockIdx.x * cap + l] = h[l];
        if (h[l]) atomicAdd(&tot[(size_t)(blockIdx.x & (TOT_COPIES - 1u)) * cap + l], h[l]);
    }
}
// one bit per slot: its agent has its home or its work in another region (the only ones a commute exchange can move)
__global__ void __launch_bounds__(256) k_mark_foreign(Params P, const uint32_t* __restrict__ st, const uint32_t* __restrict__ reg, uint32_t* __restrict__ foreign) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool f = false;
    if (i < P.n && (st[i] & ST_STATE_MASK) != ST_ABSENT) {
        const uint32_t r = reg[i], self = (uint32_t)P.region;
        f = (r & 0xFFu) != self || ((r >> 8) & 0xFFu) != self;
    }
    const unsigned b = __ballot_sync(0xFFFFFFFFu, f);
    if ((threadIdx.x & 31u) == 0 && i < P.n) foreign[i >> 5] = b;
}

// the agent leaves: record written, cell vacated, slot emptied and pushed, Counts decremented (decrement_counts,
// allocation_map.rs:291-301), occupancies of its house / office decremented (remove_migrators, :165-192)
__device__ __forceinline__ void travel_leave(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, TravelRecord* __restrict__ send, uint32_t stride,
                                             uint32_t i, uint32_t dest, uint32_t j, uint32_t p, uint32_t free_top) {
    const uint32_t s = D.st[i];
    TravelRecord r;
    r.st = s; r.t0 = D.t0[i]; r.home = D.home[i]; r.work = D.work[i]; r.reg = D.reg[i]; r.slot = i; r.from = (uint32_t)P.region; r.pad = 0;
    if (j + 1u < stride) send[(size_t)dest * stride + 1u + j] = r;
    D.grid[P.cell_offset(D.cell[i])] = 0;
    D.st[i] = ST_ABSENT;
    D.prop[i] = 0;
    if ((r.reg & 0xFFu) != (uint32_t)P.region || ((r.reg >> 8) & 0xFFu) != (uint32_t)P.region) atomicAnd(&T.foreign[i >> 5], ~(1u << (i & 31u)));
    atomicSub(D.tot + (i & (TOT_COPIES - 1u)) * 8u + count_category(s), 1u);
    // free_slots.push_back: migrators in send order (remove_migrators walks the per-region lists), commuters in selection order
    T.free_stack[free_top + p] = i;
    if (A.kind == TRAVEL_MIGRATE) {
        const uint32_t hr = house_rank(P, r.home);
        const uint32_t old = atomicSub(&T.occ_house[hr], 1u);
        if (old == OCC_ABSENT || old == 0u) atomicOr(&T.tv->err, TERR_NO_HOUSE);  // "Could not find house"
        else heap_level_moved(T.bh_house, T.tot_house, HOUSE_CAP, hr, old, old - 1u);
        if (((s >> ST_WS_SHIFT) & 3u) != WS_NA) {
            const uint32_t orank = office_rank(P, r.work);
            const uint32_t oldo = atomicSub(&T.occ_office[orank], 1u);
            if (oldo == OCC_ABSENT || oldo == 0u) atomicOr(&T.tv->err, TERR_NO_OFFICE);
            else heap_level_moved(T.bh_office, T.tot_office, OFFICE_CAP, orank, oldo, oldo - 1u);
        }
    }
}


// Peer transport (multi.cpp): where the segments go once they are packed -- rank p's receive area and flag words as seen from here
struct PeerPush {
    TravelRecord* const* peer_recv;  // [n_regions], or nullptr: the caller moves the send buffer itself
    uint32_t* const* peer_flags;
    uint32_t exchange_no;
};

__device__ __forceinline__ void leave_body(cg::grid_group& grid, const Params& P, const DevPtrs& D, TravelArgs A, const TravelPtrs& T, uint32_t* __restrict__ block_counts,
                                           uint32_t* __restrict__ ballots, uint32_t n_blocks, TravelRecord* __restrict__ send, uint32_t stride, const PeerPush& X) {
    __shared__ uint32_t s_count, s_hist[TRAVEL_MAX_REGIONS], s_base[TRAVEL_MAX_REGIONS], warp_sums[32];
    __shared__ uint16_t s_wcnt[8][TRAVEL_MAX_REGIONS];
    __shared__ unsigned long long s_thr;
    TravelVars* tv = T.tv;
    const uint32_t R = (uint32_t)T.n_regions, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t free_top = tv->free_top;
    if (blockIdx.x == 0 && tid == 0) trace_stamp(D.trace, 2, A.hour);

    // ---- 1: flags ----
    if (tid == 0) {
        unsigned long long thr = 0;
        if (A.kind == TRAVEL_MIGRATE) {
            // EngineMigrationPlan::percent_outgoing (engine_migration_plan.rs:44-49) of the region's CURRENT population; gen_bool(p)
            // panics for p > 1 (rand 0.8 Bernoulli::new; allocation_map.rs:110): an error here
            unsigned long long planned = 0;
            for (uint32_t to = 0; to < R; ++to) planned += T.plan_row[to];
            const uint32_t pop = tv->population;
            if (planned > pop) { if (blockIdx.x == 0) atomicOr(&tv->err, TERR_PERCENT); }
            else if (planned != 0 && pop != 0) thr = bernoulli_threshold((double)planned / (double)pop);
        }
        s_thr = thr;
    }
    for (uint32_t d = tid; d < R; d += blockDim.x) s_hist[d] = 0;
    __syncthreads();
    A.thr_outgoing = s_thr;
    // every grid block owns a contiguous slice of the blocks of 1024 slots: it flags them (phase 1), turns their counts into offsets
    // inside the slice right away, and walks them again in phase 3
    const uint32_t sb_per = (n_blocks + gridDim.x - 1u) / gridDim.x;
    const uint32_t sb0 = min(n_blocks, blockIdx.x * sb_per), sb1 = min(n_blocks, sb0 + sb_per);
    if (A.kind == TRAVEL_COMMUTE) {
        // only an agent whose home or work is in another region can commute: one bit per slot says so (T.foreign), so the flags of a
        // block of 1024 slots are a filter of its 32 bitmap words -- one warp per block, lane l owns word l
        for (uint32_t g = sb0 + warp; g < sb1; g += 8u) {
            const uint32_t wi = g * 32u + lane;
            uint32_t bits = wi * 32u < P.n ? T.foreign[wi] : 0u, out = 0;
            while (bits) {
                const uint32_t bit = (uint32_t)__ffs((int)bits) - 1u, i = wi * 32u + bit;
                bits &= bits - 1u;
                const uint32_t f = i < P.n ? travel_flag(P, A, i, D.st[i], D.reg[i]) : 0u;
                if (f) {
                    out |= 1u << bit;
                    if (f - 1u < R) atomicAdd(&s_hist[f - 1u], 1u);
                    else atomicOr(&tv->err, TERR_BAD_REGION);
                }
            }
            ballots[wi] = out;
            uint32_t c = (uint32_t)__popc(out);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
            if (lane == 0) block_counts[g] = c;
        }
        __syncthreads();
    }
    for (uint32_t sb = sb0; sb < sb1 && A.kind != TRAVEL_COMMUTE; ++sb) {  // 256 threads x 4 consecutive slots; eight threads assemble one 32-slot ballot word
        if (tid == 0) s_count = 0;
        __syncthreads();
        const uint32_t i0 = (sb * 256u + tid) * 4u;
        uint32_t f[4] = {0, 0, 0, 0};
        if (i0 + 3u < P.n) {
            const uint4 s4 = __ldcs(reinterpret_cast<const uint4*>(D.st + i0)), r4 = __ldcs(reinterpret_cast<const uint4*>(D.reg + i0));
            f[0] = travel_flag(P, A, i0, s4.x, r4.x); f[1] = travel_flag(P, A, i0 + 1u, s4.y, r4.y);
            f[2] = travel_flag(P, A, i0 + 2u, s4.z, r4.z); f[3] = travel_flag(P, A, i0 + 3u, s4.w, r4.w);
        } else {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (i0 + j < P.n) f[j] = travel_flag(P, A, i0 + j, D.st[i0 + j], D.reg[i0 + j]);
        }
        uint32_t nib = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (f[j]) {
                nib |= 1u << j;
                if (A.kind == TRAVEL_COMMUTE) {
                    if (f[j] - 1u < R) atomicAdd(&s_hist[f[j] - 1u], 1u);
                    else atomicOr(&tv->err, TERR_BAD_REGION);
                }
            }
        uint32_t w = nib << ((tid & 7u) * 4u);
        w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
        w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
        w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
        if ((tid & 7u) == 0) ballots[i0 >> 5] = w;  // written for every group of the block, also beyond P.n
        uint32_t c = (tid & 7u) == 0 ? (uint32_t)__popc(w) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        if (lane == 0 && c) atomicAdd(&s_count, c);
        __syncthreads();
        if (tid == 0) block_counts[sb] = s_count;
    }
    if (A.kind == TRAVEL_COMMUTE)
        for (uint32_t d = tid; d < R; d += blockDim.x)
            if (s_hist[d]) atomicAdd(&tv->hist[d], s_hist[d]);
    __syncthreads();
    {   // offsets inside the slice; the slice's total for block 0
        uint32_t carry = 0;
        for (uint32_t base = sb0; base < sb1; base += blockDim.x) {
            const uint32_t b = base + tid;
            const uint32_t v = b < sb1 ? block_counts[b] : 0u;
            uint32_t sum;
            const uint32_t ex = block_exclusive_scan(v, warp_sums, &sum);
            if (b < sb1) block_counts[b] = carry + ex;
            carry += sum;
        }
        if (tid == 0) T.hslice[blockIdx.x] = carry;
    }
    grid.sync();

    // ---- 2: offsets and the plan.  Migrators: EngineMigrationPlan::alloc_outgoing_to_regions (engine_migration_plan.rs:51-77) +
    // MigratorsByRegion::alloc_citizens (migrators_by_engine.rs:34-55): regions in plan order take floor(share * total) from the
    // front of the list, the rest stays.  Commuters: CommutersByRegion::get_commuters_by_region (commuters_by_region.rs:59-78). ----
    if (blockIdx.x == 0) {
        uint32_t mine = 0;
        for (uint32_t g = tid; g < gridDim.x; g += blockDim.x) mine += T.hslice[g];
        uint32_t total;
        block_exclusive_scan(mine, warp_sums, &total);
        if (tid == 0) {
            tv->total = total;
            uint32_t front = 0;
            if (A.kind == TRAVEL_MIGRATE) {
                uint64_t planned_total = 0;
                for (uint32_t to = 0; to < R; ++to) planned_total += T.plan_row[to];
                for (uint32_t to = 0; to < R; ++to) {
                    uint32_t count = 0;
                    if ((int)to != P.region && T.plan_row[to] != 0) {
                        const double share = (double)T.plan_row[to] / (double)planned_total;
                        count = (uint32_t)(int32_t)(share * (double)(int32_t)total);
                        if (count > total - front) count = total - front;
                    }
                    tv->cnt[to] = count; tv->base[to] = front;
                    front += count;
                }
            } else {
                for (uint32_t to = 0; to < R; ++to) { tv->cnt[to] = tv->hist[to]; tv->base[to] = front; front += tv->hist[to]; tv->hist[to] = 0; }
                if (total > T.list_cap) atomicOr(&tv->err, TERR_LIST_OVERFLOW);
            }
            tv->n_send = front;
        }
        __syncthreads();
        for (uint32_t to = tid; to < R; to += blockDim.x) {
            if (tv->cnt[to] + 1u > (T.seg_cap ? min(T.seg_cap[to], stride) : stride)) atomicOr(&tv->err, TERR_SEGMENT_OVERFLOW);
            TravelRecord h{};
            h.st = min(tv->cnt[to], stride - 1u);  // header: records in this segment
            h.from = (uint32_t)P.region;
            send[(size_t)to * stride] = h;
        }
        // An overflow (or any earlier error of the run) is known by now: then nobody is removed -- the region stays intact, the host
        // reports the error -- and the headers promise no records.
        __syncthreads();
        if (tid == 0) tv->abort = tv->err;
        __syncthreads();
        if (tv->err) {
            for (uint32_t to = tid; to < R; to += blockDim.x) send[(size_t)to * stride].st = 0;
            if (tid == 0) tv->n_send = 0;
        }
    }
    grid.sync();
    const bool abort = tv->abort != 0;
    const uint32_t n_send = tv->n_send, total = tv->total;
    uint32_t slice_off;  // leavers in the slices before this block's
    {
        uint32_t before = 0;
        for (uint32_t g = tid; g < blockIdx.x; g += blockDim.x) before += T.hslice[g];
        block_exclusive_scan(before, warp_sums, &slice_off);
    }

    // ---- 3: the leavers in slot order: one warp per block of 1024 slots, lane l owns ballot word l ----
    if (!abort)
        for (uint32_t g = sb0 + warp; g < sb1; g += 8u) {
            uint32_t b = ballots[g * 32u + lane];
            const uint32_t c = (uint32_t)__popc(b), incl = warp_inclusive(c);
            if (__shfl_sync(0xFFFFFFFFu, incl, 31) == 0) continue;
            uint32_t p = slice_off + block_counts[g] + incl - c;
            while (b) {
                const uint32_t i = (g * 32u + lane) * 32u + (uint32_t)__ffs((int)b) - 1u;
                b &= b - 1u;
                if (A.kind == TRAVEL_COMMUTE) {  // Citizen::is_commuter: the work region in the morning, the home region in the evening
                    const uint32_t reg = D.reg[i];
                    T.list_slot[p] = i;
                    T.list_dest[p] = A.hour_of_day == 7u ? (reg >> 8) & 0xFFu : reg & 0xFFu;
                } else if (p < n_send) {  // surplus candidates stay
                    uint32_t dest = 0;
                    while (!(p >= tv->base[dest] && p < tv->base[dest] + tv->cnt[dest])) ++dest;
                    travel_leave(P, D, A, T, send, stride, i, dest, p - tv->base[dest], p, free_top);
                }
                ++p;
            }
        }
    if (A.kind == TRAVEL_COMMUTE && !abort) {
        const uint32_t n_chunks = (total + 255u) / 256u;
        grid.sync();
        // ---- 4: leavers per destination in every chunk of 256 list entries ----
        for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            for (uint32_t d = tid; d < R; d += blockDim.x) s_hist[d] = 0;
            __syncthreads();
            const uint32_t p = c * 256u + tid;
            if (p < total) atomicAdd(&s_hist[T.list_dest[p]], 1u);
            __syncthreads();
            for (uint32_t d = tid; d < R; d += blockDim.x) T.chunk_dest[(size_t)c * R + d] = s_hist[d];
            __syncthreads();
        }
        grid.sync();
        // ---- 5: stable index inside the destination's segment; the agent leaves ----
        for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            for (uint32_t d = tid; d < R; d += blockDim.x) {
                uint32_t before = 0;
                for (uint32_t q = 0; q < c; ++q) before += T.chunk_dest[(size_t)q * R + d];
                s_base[d] = before;
            }
            for (uint32_t k = tid; k < 8u * R; k += blockDim.x) s_wcnt[k / R][k % R] = 0;
            __syncthreads();
            const uint32_t p = c * 256u + tid;
            const bool active = p < total;
            const uint32_t d = active ? T.list_dest[p] : 0u;
            const unsigned same = __match_any_sync(0xFFFFFFFFu, active ? d : (0x100u | lane));
            const uint32_t in_warp = (uint32_t)__popc(same & ((1u << lane) - 1u));
            if (active && in_warp == 0) s_wcnt[warp][d] = (uint16_t)__popc(same);
            __syncthreads();
            if (active) {
                uint32_t rank = s_base[d] + in_warp;
                for (uint32_t w2 = 0; w2 < warp; ++w2) rank += s_wcnt[w2][d];
                travel_leave(P, D, A, T, send, stride, T.list_slot[p], d, rank, p, free_top);
            }
            __syncthreads();
        }
    }
    grid.sync();
    // ---- 6 ----
    if (blockIdx.x == 0 && tid == 0) trace_stamp(D.trace, 3, A.hour);
    if (blockIdx.x == 0 && tid == 0 && !abort) {  // the free-slot stack grows by the leavers
        tv->free_top = free_top + n_send;
        tv->population -= n_send;
    }
    // ---- 7: peer transport: the used part of every destination's segment goes into that rank's receive area (peer memory over
    // NVLink / NVSwitch, or this rank's own for p == me); then its flag there says so: flags[me] = the exchange's number ----
    if (X.peer_recv) {
        const uint32_t gtid = blockIdx.x * blockDim.x + tid, n_threads = gridDim.x * blockDim.x;
        for (uint32_t p = 0; p < R; ++p) {  // every thread of the grid copies: the stores to a peer are in flight together
            const uint32_t n = min(send[(size_t)p * stride].st, stride - 1u) + 1u;  // header + records
            const uint4* src = reinterpret_cast<const uint4*>(send + (size_t)p * stride);
            uint4* dst = reinterpret_cast<uint4*>(X.peer_recv[p] + ((size_t)(X.exchange_no & 1u) * R + (uint32_t)P.region) * stride);
            for (uint32_t k = gtid; k < 2u * n; k += n_threads) dst[k] = src[k];
        }
        __threadfence_system();
        grid.sync();
        if (gtid < R) {
            __threadfence_system();
            *(volatile uint32_t*)(X.peer_flags[gtid] + P.region) = X.exchange_no;
        }
    }
}

__global__ void __launch_bounds__(256) k_travel_leave(Params P, DevPtrs D, TravelArgs A, TravelPtrs T, uint32_t* __restrict__ block_counts, uint32_t* __restrict__ ballots,
                                                       uint32_t n_blocks, TravelRecord* __restrict__ send, uint32_t stride, PeerPush X) {
    cg::grid_group grid = cg::this_grid();
    leave_body(grid, P, D, A, T, block_counts, ballots, n_blocks, send, stride, X);
}

// ---- arriving -----------------------------------------------------------------------------------------------------------
// One occupancy heap (houses or offices) and the arrivals it serves.  "Water filling": the pop sequence of the reference's
// BinaryHeap of areas for K arrivals, in parallel (see the head of this file).
struct HeapJob {
    uint32_t* occ;             // occupants per area in tie order, OCC_ABSENT = not in the heap
    uint32_t n, n_blocks, cap; // areas, blocks of 256 areas, levels (HOUSE_CAP / OFFICE_CAP)
    uint32_t *bh, *pref;       // per-block level histograms [n_blocks][cap]; per-level block prefixes [cap][n_blocks]
    uint32_t *tot, *slice;     // areas per level overall [cap]; scratch: per grid block one slice total
    FillPlan* plan;
    const uint32_t* k_src;     // number of pops (device-side count)
    uint32_t err_bit;
    uint32_t* out;             // per arrival: the area it got (index in tie order)
    const uint32_t* pop_index; // which pop serves arrival k (nullptr: k)
};
struct HeapJobs {
    HeapJob job[2];
    int n_jobs;
};

// pop p of heap `job` -> the area it returns (index in tie order).  One warp.
__device__ __forceinline__ uint32_t heap_pop_area(const HeapJob& job, uint32_t p, uint32_t lane) {
    const FillPlan* plan = job.plan;
    if (p >= plan->K) return 0xFFFFFFFFu;
    const uint32_t n_blocks = job.n_blocks, n = job.n;
    uint32_t l = plan->L;
    while (l < plan->l_last && p >= plan->start[l + 1]) ++l;
    const uint32_t q = p - plan->start[l];
    const uint32_t* pl = job.pref + (size_t)l * n_blocks;
    uint32_t lo = 0, hi = n_blocks;  // largest b with pl[b] <= q (pl[0] == 0): 32-ary search, the lanes probe together
    while (hi - lo > 1u) {
        const uint32_t step = (hi - lo + 31u) / 32u, idx = lo + lane * step;
        const unsigned ok = __ballot_sync(0xFFFFFFFFu, idx < hi && pl[idx] <= q);
        const uint32_t m = 31u - (uint32_t)__clz((int)ok);  // lane 0 always qualifies
        lo += m * step;
        hi = min(hi, lo + step);
    }
    uint32_t left = q - pl[lo];  // the left-th area with occupancy <= l inside block lo (256 areas)
    uint32_t found = 0xFFFFFFFFu;
    uint32_t o[8];
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c) {
        const uint32_t r = lo * 256u + c * 32u + lane;
        o[c] = r < n ? job.occ[r] : OCC_ABSENT;
    }
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, o[c] <= l);
        const uint32_t cnt = (uint32_t)__popc(m);
        if (found == 0xFFFFFFFFu) {
            if (left < cnt) {
                unsigned mm = m;  // position of the left-th set bit of m
                for (uint32_t t = 0; t < left; ++t) mm &= mm - 1u;
                found = lo * 256u + c * 32u + (uint32_t)(__ffs(mm) - 1);
            } else left -= cnt;
        }
    }
    return found;
}

// arrival k -> the k-th slot from the top of the free stack.  Citizen::from_migrator / from_commuter (citizen/mod.rs:113-154):
// immunity, vaccinated, uses_public_transport and the disease state travel; hospitalized, isolated, work_quarantined reset;
// work status becomes NA (migrator) or Normal (commuter); current_area = the housing strip.  The cell is assigned by the
// placement rounds.
__device__ __forceinline__ void travel_install(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, uint32_t k, uint32_t free_top) {
    const TravelRecord r = T.arrivals[k];
    const uint32_t i = T.free_stack[free_top - 1u - k];
    const uint32_t keep = ST_STATE_MASK | (3u << ST_SEV_SHIFT) | (7u << ST_IMM_SHIFT) | ST_VACC | ST_PT | (ST_DAY_MAX << ST_DAY_SHIFT);
    const uint32_t ws = A.kind == TRAVEL_MIGRATE ? WS_NA : WS_NORMAL;
    const uint32_t s = (r.st & keep) | (ws << ST_WS_SHIFT) | (AK_HOUSING << ST_AREA_SHIFT);
    const uint32_t self = (uint32_t)P.region;
    D.st[i] = s;
    D.t0[i] = r.t0;
    D.wsa[i] = 0;
    D.prop[i] = 0;
    if (A.kind == TRAVEL_MIGRATE) {
        const uint32_t house = T.arr_house[k];
        D.home[i] = house != 0xFFFFFFFFu ? house_origin_of_rank(P, house) : 0u;
        D.work[i] = 0;  // WorkStatus::NA after from_migrator: the office only counts in the occupancy heap
        D.reg[i] = self | (self << 8);
    } else {
        D.home[i] = r.home;
        const bool assign_office = A.hour == 7u;  // sic: the absolute hour (allocation_map.rs:260), i.e. the first day only
        const uint32_t office = assign_office ? T.arr_office[k] : 0xFFFFFFFFu;
        D.work[i] = office != 0xFFFFFFFFu ? office_origin_of_rank(P, office) : r.work;
        D.reg[i] = (r.reg & 0xFFu) | ((assign_office ? self : ((r.reg >> 8) & 0xFFu)) << 8);
    }
    if ((D.reg[i] & 0xFFu) != self || ((D.reg[i] >> 8) & 0xFFu) != self) atomicOr(&T.foreign[i >> 5], 1u << (i & 31u));
    atomicAdd(D.tot + (i & (TOT_COPIES - 1u)) * 8u + count_category(s), 1u);
}


// ---- placement rounds (select_starting_points, allocation_map.rs:339-347) ------------------------------------------
// In round a every still-unplaced arrival k walks its own candidate sequence Philox(seed, k, hour, DOM_ARRIVAL) blocks
// a * PLACE_TRIES .. a * PLACE_TRIES + PLACE_TRIES - 1 (x in [sx, ex), y in [sy, ey)) and proposes the FIRST candidate that is
// vacant now (start-of-hour occupants and earlier rounds' winners count as occupied); of several proposals for one cell the
// lowest k wins; a loser, or an arrival whose candidates were all occupied, tries again in the next round.
constexpr uint32_t PLACE_TRIES = 8, PLACE_MAX_ROUNDS = 64;
__device__ __forceinline__ uint32_t arrival_cell(const Params& P, const TravelArgs& A, uint32_t k, uint32_t block) {
    const U4 o = philox4x32_10(k, A.hour, block, DOM_ARRIVAL, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    const Rect& r = A.kind == TRAVEL_MIGRATE ? P.housing() : P.transport();
    const uint32_t x = (uint32_t)r.sx + __umulhi(o.x, (uint32_t)(r.ex - r.sx));
    const uint32_t y = (uint32_t)r.sy + __umulhi(o.y, (uint32_t)(r.ey - r.sy));
    return (y << CELL_BITS) | x;
}
__device__ __forceinline__ uint32_t first_vacant_candidate(const Params& P, const DevPtrs& D, const TravelArgs& A, uint32_t k, uint32_t attempt) {
    for (uint32_t t = 0; t < PLACE_TRIES; ++t) {
        const uint32_t c = arrival_cell(P, A, k, attempt * PLACE_TRIES + t);
        if (D.grid[P.cell_offset(c)] == 0) return c;
    }
    return 0xFFFFFFFFu;
}
__device__ __forceinline__ uint32_t hash_cell(uint32_t c) { return (c * 0x9E3779B1u) >> 7; }


// All of an exchange's arrivals in ONE cooperative launch (k_travel_arrive); grid-wide barriers separate its phases:
//   1  recv: one segment per source region (header + records) -> arrivals[k], k ascending in (source region, index)
//   2  migrators: rank of every arrival among the working ones (they pop an office each as well as a house)
//   3  water filling of the occupancy heaps, when houses / offices are assigned: level histograms, the plan of the K pops, block
//      prefixes of the levels the plan uses, the area of every pop, the areas' new occupancies
//   4  install: arrival k -> the k-th slot from the top of the free stack (Citizen::from_migrator / from_commuter)
//   5  placement rounds until everybody stands on a cell of its own (the hash table of a round's proposals is emptied by its
//      proposers)
//   6  what used to need the host: the arrivals' slots leave the free stack, the population grows, and the Counts row of the
//      exchange hour -- Counts after remove_* / assimilate_* adjusted them (epidemiology_simulation.rs:467-492) -- goes to the
//      counts ring with the population and the error flags next to it
struct PeerWait {
    const uint32_t* flags;  // [n] written by the peers (k_travel_push), or nullptr: the receive buffer is complete at launch
    uint32_t exchange_no;
    int n;
};

__device__ __forceinline__ void arrive_body(cg::grid_group& grid, const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T,
                                            const TravelRecord* __restrict__ recv, uint32_t stride, const HeapJobs& J, const PeerWait& W) {
    __shared__ uint32_t s_base[TRAVEL_MAX_REGIONS + 1], s_h[OFFICE_CAP], warp_sums[32];
    TravelVars* tv = T.tv;
    const uint32_t R = (uint32_t)T.n_regions, free_top = tv->free_top;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, gtid = blockIdx.x * blockDim.x + tid, n_threads = gridDim.x * blockDim.x;

    if (gtid == 0) trace_stamp(D.trace, 4, A.hour);
    // ---- 0: peer transport: every source's segment must have landed (its flag carries this exchange's number) ----
    if (W.flags && tid == 0) {
        const long long t0 = clock64();
        for (int q = 0; q < W.n; ++q)
            while (*(volatile const uint32_t*)(W.flags + q) < W.exchange_no) {
                if (clock64() - t0 > 60ll * 2000000000ll) {  // ~60 s: a peer died; report instead of hanging the device
                    if (blockIdx.x == 0) atomicOr(&tv->err, TERR_PEER_TIMEOUT);
                    break;
                }
                __nanosleep(200);
            }
        __threadfence_system();
    }
    if (gtid == 0) trace_stamp(D.trace, 5, A.hour);
    // ---- 1: gather ----
    if (tid == 0) {
        uint32_t n = 0;
        for (uint32_t q = 0; q < R; ++q) { s_base[q] = n; n += min(__ldcg(&recv[(size_t)q * stride].st), stride - 1u); }
        s_base[R] = n;
    }
    __syncthreads();
    uint32_t n_in = s_base[R];
    {
        uint32_t err = 0;
        if (n_in > T.list_cap) { err |= TERR_LIST_OVERFLOW; n_in = 0; }
        if (n_in > free_top) { err |= TERR_NO_SLOTS; n_in = 0; }  // "region is out of agent slots"
        if (gtid == 0) {
            if (err) atomicOr(&tv->err, err);
            for (uint32_t q = 0; q < R; ++q) tv->cnt[q] = s_base[q + 1] - s_base[q];  // the headers, for the host
            tv->n_in = n_in;
            tv->n_working = 0;
            tv->pend2[0] = tv->pend2[1] = 0;
        }
    }
    for (uint32_t k = gtid; k < n_in; k += n_threads) {
        uint32_t s = 0;
        while (k >= s_base[s + 1]) ++s;
        const uint4* src = reinterpret_cast<const uint4*>(recv + (size_t)s * stride + 1u + (k - s_base[s]));  // L2 loads: a peer may have written this
        uint4* dst = reinterpret_cast<uint4*>(T.arrivals + k);
        dst[0] = __ldcg(src);
        dst[1] = __ldcg(src + 1);
        T.placed[k] = 0;
    }
    grid.sync();

    // ---- 2: migrators: arr_widx[k] = number of working arrivals before k (assimilate_migrators, allocation_map.rs:214-243): every grid
    // block counts its slice of the arrivals, then ranks it behind the slices before it ----
    if (A.kind == TRAVEL_MIGRATE) {
        const uint32_t per = (n_in + gridDim.x - 1u) / gridDim.x;
        const uint32_t k0 = min(n_in, blockIdx.x * per), k1 = min(n_in, k0 + per);
        uint32_t mine = 0;
        for (uint32_t k = k0 + tid; k < k1; k += blockDim.x) mine += ((T.arrivals[k].st >> ST_WS_SHIFT) & 3u) != WS_NA;
        uint32_t slice_total;
        block_exclusive_scan(mine, warp_sums, &slice_total);
        if (tid == 0) T.hslice[blockIdx.x] = slice_total;
        grid.sync();
        uint32_t before = 0, all = 0;
        for (uint32_t g = tid; g < gridDim.x; g += blockDim.x) { const uint32_t v = T.hslice[g]; all += v; if (g < blockIdx.x) before += v; }
        uint32_t carry, total;
        block_exclusive_scan(before, warp_sums, &carry);
        block_exclusive_scan(all, warp_sums, &total);
        for (uint32_t base = k0; base < k1; base += blockDim.x) {
            const uint32_t k = base + tid;
            const uint32_t f = (k < k1 && ((T.arrivals[k].st >> ST_WS_SHIFT) & 3u) != WS_NA) ? 1u : 0u;
            uint32_t sum;
            const uint32_t ex = block_exclusive_scan(f, warp_sums, &sum);
            if (k < k1) T.arr_widx[k] = f ? carry + ex : 0xFFFFFFFFu;
            carry += sum;
        }
        if (gtid == 0) tv->n_working = total;
        grid.sync();
    }

    // ---- 3: water filling ----
    if (J.n_jobs > 0 && n_in != 0) {
        // the level structure of the K pops: grid block j plans heap j
        if ((int)blockIdx.x < J.n_jobs) {
            const HeapJob& job = J.job[blockIdx.x];
            const uint32_t CAP = job.cap;
            for (uint32_t l = tid; l < CAP; l += blockDim.x) {  // areas per occupancy level (kept current, heap_level_moved)
                uint32_t sum = 0;
                for (uint32_t c = 0; c < TOT_COPIES; ++c) sum += job.tot[(size_t)c * CAP + l];
                s_h[l] = sum;
            }
            __syncthreads();
            if (tid == 0) {
                FillPlan* plan = job.plan;
                const uint32_t K = *job.k_src;
                plan->K = K; plan->L = 0; plan->l_last = 0; plan->last_count = 0;
                if (K != 0) {
                    uint32_t L = 0;
                    while (L < CAP && s_h[L] == 0) ++L;
                    uint64_t cum = 0, n_le = 0;
                    bool done = false;
                    for (uint32_t l = L; l < CAP; ++l) {
                        n_le += s_h[l];
                        plan->start[l] = (uint32_t)cum;
                        if (cum + n_le >= K) { plan->l_last = l; plan->last_count = K - (uint32_t)cum; done = true; break; }
                        cum += n_le;
                    }
                    plan->L = L;
                    if (!done) {  // "Couldn't find any house / offices with free space!" (allocation_map.rs:222, :254)
                        plan->K = 0;
                        atomicOr(&tv->err, job.err_bit);
                    }
                }
            }
        }
        grid.sync();
        // pref[l * n_blocks + b] = number of areas with occupancy <= l in area blocks before b, for the levels the plan uses: every
        // grid block scans its slice of the area blocks (all levels), then adds the totals of the slices before it
        for (int j = 0; j < J.n_jobs; ++j) {
            const HeapJob& job = J.job[j];
            const FillPlan* plan = job.plan;
            if (plan->K == 0) continue;
            const uint32_t CAP = job.cap, L = plan->L, l_last = plan->l_last, nb = job.n_blocks;
            const uint32_t per = (nb + gridDim.x - 1u) / gridDim.x;  // area blocks per grid block
            const uint32_t b0 = min(nb, blockIdx.x * per), b1 = min(nb, b0 + per);
            for (uint32_t l = L; l <= l_last; ++l) {
                uint32_t carry = 0;
                for (uint32_t base = b0; base < b1; base += blockDim.x) {
                    const uint32_t b = base + tid;
                    uint32_t v = 0;
                    if (b < b1)
                        for (uint32_t q = L; q <= l; ++q) v += job.bh[(size_t)b * CAP + q];
                    uint32_t sum;
                    const uint32_t ex = block_exclusive_scan(v, warp_sums, &sum);
                    if (b < b1) job.pref[(size_t)l * nb + b] = carry + ex;
                    carry += sum;
                }
                if (tid == 0) job.slice[(size_t)(l - L) * gridDim.x + blockIdx.x] = carry;
            }
        }
        grid.sync();
        for (int j = 0; j < J.n_jobs; ++j) {
            const HeapJob& job = J.job[j];
            const FillPlan* plan = job.plan;
            if (plan->K == 0) continue;
            const uint32_t L = plan->L, l_last = plan->l_last, nb = job.n_blocks;
            const uint32_t per = (nb + gridDim.x - 1u) / gridDim.x;
            const uint32_t b0 = min(nb, blockIdx.x * per), b1 = min(nb, b0 + per);
            if (b0 >= b1) continue;
            for (uint32_t l = L; l <= l_last; ++l) {
                uint32_t before = 0;
                for (uint32_t g = tid; g < blockIdx.x; g += blockDim.x) before += job.slice[(size_t)(l - L) * gridDim.x + g];
                uint32_t total;
                block_exclusive_scan(before, warp_sums, &total);
                for (uint32_t b = b0 + tid; b < b1; b += blockDim.x) job.pref[(size_t)l * nb + b] += total;
            }
        }
        grid.sync();
        // pop p -> the area it returns: one warp per arrival and heap
        {
            const uint32_t gw = gtid >> 5, n_warps = n_threads >> 5;
            for (int j = 0; j < J.n_jobs; ++j) {
                const HeapJob& job = J.job[j];
                for (uint32_t k = gw; k < n_in; k += n_warps) {
                    const uint32_t p = job.pop_index ? job.pop_index[k] : k;  // which pop of this heap serves arrival k (0xFFFFFFFF: none)
                    const uint32_t area = heap_pop_area(job, p, lane);
                    if (lane == 0) job.out[k] = area;
                }
            }
        }
        grid.sync();
        // every area that was popped comes back with one more occupant per pop
        for (int j = 0; j < J.n_jobs; ++j) {
            const HeapJob& job = J.job[j];
            for (uint32_t k = gtid; k < n_in; k += n_threads) {
                const uint32_t area = job.out[k];
                if (area == 0xFFFFFFFFu) continue;
                const uint32_t old = atomicAdd(&job.occ[area], 1u);
                heap_level_moved(job.bh, job.tot, job.cap, area, old, old + 1u);
            }
        }
        grid.sync();
    }

    // ---- 4: install ----
    for (uint32_t k = gtid; k < n_in; k += n_threads) travel_install(P, D, A, T, k, free_top);
    grid.sync();

    // ---- 5: placement rounds ----
    uint32_t attempt = 0;
    if (n_in)
        for (;; ++attempt) {
            uint32_t* pend = &tv->pend2[attempt & 1u];
            if (gtid == 0) tv->pend2[(attempt + 1u) & 1u] = 0;
            for (uint32_t k = gtid; k < n_in; k += n_threads) {  // propose
                if (T.placed[k]) continue;
                const uint32_t c = first_vacant_candidate(P, D, A, k, attempt);
                T.list_pos[k] = c;
                if (c == 0xFFFFFFFFu) continue;
                uint32_t slot = hash_cell(c) & T.table_mask;
                for (;;) {
                    const uint32_t old = atomicCAS(&T.table_keys[slot], 0u, c + 1u);
                    if (old == 0u || old == c + 1u) { atomicMin(&T.table_vals[slot], k); break; }
                    slot = (slot + 1u) & T.table_mask;
                }
            }
            grid.sync();
            for (uint32_t k = gtid; k < n_in; k += n_threads) {  // grant: the lowest arrival among a cell's proposers takes it
                if (T.placed[k]) continue;
                const uint32_t c = T.list_pos[k];
                bool won = false;
                if (c != 0xFFFFFFFFu) {
                    uint32_t slot = hash_cell(c) & T.table_mask;
                    while (T.table_keys[slot] != c + 1u) slot = (slot + 1u) & T.table_mask;
                    won = T.table_vals[slot] == k;
                    T.list_dest[k] = slot;
                }
                if (won) {
                    const uint32_t i = T.free_stack[free_top - 1u - k];
                    D.cell[i] = c;
                    D.grid[P.cell_offset(c)] = (uint8_t)cell_byte(P, D.st[i]);
                    T.placed[k] = 2;  // placed in this round: its table entry is still to be emptied
                } else atomicAdd(pend, 1u);
            }
            grid.sync();
            for (uint32_t k = gtid; k < n_in; k += n_threads) {  // empty the table
                const uint8_t pl = T.placed[k];
                if (pl == 1) continue;
                if (pl == 2) T.placed[k] = 1;
                if (T.list_pos[k] == 0xFFFFFFFFu) continue;
                const uint32_t slot = T.list_dest[k];
                T.table_keys[slot] = 0u;
                T.table_vals[slot] = 0xFFFFFFFFu;
            }
            const uint32_t left = *pend;
            grid.sync();
            if (left == 0) break;
            if (attempt + 1u >= PLACE_MAX_ROUNDS) {  // "Not enough locations are available for travellers"
                if (gtid == 0) atomicOr(&tv->err, TERR_NO_PLACE);
                break;
            }
        }

    // ---- 6 ----
    if (gtid != 0) return;
    trace_stamp(D.trace, 6, A.hour);
    tv->pending = 0;
    tv->rounds = max(tv->rounds, n_in ? attempt + 1u : 0u);
    tv->free_top = free_top - n_in;  // the arrivals' slots leave the free stack
    tv->population += n_in;
    if (A.row_index == 0xFFFFFFFFu) return;
    uint32_t* row = D.counts + (size_t)A.row_index * 8;
    for (uint32_t c = 0; c < 6; ++c) {
        uint32_t v = 0;
        for (uint32_t k = 0; k < TOT_COPIES; ++k) v += D.tot[k * 8u + c];
        row[c] = v;
    }
    row[6] = tv->population;
    row[7] = tv->err;
}

__global__ void __launch_bounds__(256) k_travel_arrive(Params P, DevPtrs D, TravelArgs A, TravelPtrs T, const TravelRecord* __restrict__ recv, uint32_t stride, HeapJobs J,
                                                        PeerWait W) {
    cg::grid_group grid = cg::this_grid();
    arrive_body(grid, P, D, A, T, recv, stride, J, W);
}

// Peer transport: the whole exchange of a region in one launch -- leave, push the segments into the peers' memory, wait for theirs,
// arrive (no kernel boundary, no launch latency between the two halves)
__global__ void __launch_bounds__(256) k_travel_exchange(Params P, DevPtrs D, TravelArgs A, TravelPtrs T, uint32_t* __restrict__ block_counts, uint32_t* __restrict__ ballots,
                                                          uint32_t n_blocks, TravelRecord* __restrict__ send, uint32_t stride, PeerPush X,
                                                          const TravelRecord* __restrict__ recv, HeapJobs J, PeerWait W) {
    cg::grid_group grid = cg::this_grid();
    leave_body(grid, P, D, A, T, block_counts, ballots, n_blocks, send, stride, X);
    grid.sync();  // the stack top and the population as the leavers left them
    arrive_body(grid, P, D, A, T, recv, stride, J, W);
}

// ---- launchers ---------------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint32_t n) { return (n + 255u) / 256u; }

// block_counts: [nb + 1] counts / offsets | [nb * 32] ballots
cudaError_t launch_travel_leave(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, uint32_t* block_counts, TravelRecord* send,
                                uint32_t stride, unsigned grid_blocks, TravelRecord* const* peer_recv, uint32_t* const* peer_flags, uint32_t exchange_no, cudaStream_t s) {
    uint32_t nb = (P.n + SEL_BLOCK - 1u) / SEL_BLOCK;
    uint32_t* ballots = block_counts + nb + 1;
    PeerPush X{peer_recv, peer_flags, exchange_no};
    void* args[] = {(void*)&P, (void*)&D, (void*)&A, (void*)&T, (void*)&block_counts, (void*)&ballots, (void*)&nb, (void*)&send, (void*)&stride, (void*)&X};
    return cudaLaunchCooperativeKernel((const void*)k_travel_leave, dim3(std::min(grid_blocks, std::max(nb, 1u))), dim3(256), args, 0, s);
}

// the kernels read the number of arrivals from the segment headers
static HeapJobs heap_jobs(const TravelArgs& A, const TravelPtrs& T, uint32_t n_houses, uint32_t n_offices, unsigned grid_blocks) {
    HeapJobs J{};
    const HeapJob house{T.occ_house, n_houses, blocks_for(n_houses), HOUSE_CAP, T.bh_house, T.pref_house, T.tot_house, T.hslice, T.plan_house, &T.tv->n_in, TERR_HOUSES_FULL, T.arr_house, nullptr};
    HeapJob office{T.occ_office, n_offices, blocks_for(n_offices), OFFICE_CAP, T.bh_office, T.pref_office, T.tot_office, T.hslice + (size_t)OFFICE_CAP * grid_blocks,
                   T.plan_office, &T.tv->n_in, TERR_OFFICES_FULL, T.arr_office, nullptr};
    if (A.kind == TRAVEL_MIGRATE) {
        // assimilate_migrators (allocation_map.rs:214-243): every arrival pops a house, the working ones an office as well
        office.k_src = &T.tv->n_working;
        office.pop_index = T.arr_widx;
        J.n_jobs = 2;
        J.job[0] = house;
        J.job[1] = office;
    } else if (A.hour == 7u) {
        // assimilate_commuters (allocation_map.rs:245-277): an office is assigned at the absolute hour 7 only (:260)
        J.n_jobs = 1;
        J.job[0] = office;
    }
    return J;
}
cudaError_t launch_travel_arrive(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, const TravelRecord* recv, uint32_t stride,
                                 uint32_t n_houses, uint32_t n_offices, unsigned grid_blocks, const uint32_t* wait_flags, uint32_t exchange_no, cudaStream_t s) {
    PeerWait W{wait_flags, exchange_no, T.n_regions};
    HeapJobs J = heap_jobs(A, T, n_houses, n_offices, grid_blocks);
    void* args[] = {(void*)&P, (void*)&D, (void*)&A, (void*)&T, (void*)&recv, (void*)&stride, (void*)&J, (void*)&W};
    return cudaLaunchCooperativeKernel((const void*)k_travel_arrive, dim3(grid_blocks), dim3(256), args, 0, s);
}
cudaError_t launch_travel_exchange(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, uint32_t* block_counts, TravelRecord* send, uint32_t stride,
                                   const TravelRecord* recv, uint32_t n_houses, uint32_t n_offices, unsigned grid_blocks, TravelRecord* const* peer_recv,
                                   uint32_t* const* peer_flags, const uint32_t* wait_flags, uint32_t exchange_no, cudaStream_t s) {
    uint32_t nb = (P.n + SEL_BLOCK - 1u) / SEL_BLOCK;
    uint32_t* ballots = block_counts + nb + 1;
    PeerPush X{peer_recv, peer_flags, exchange_no};
    PeerWait W{wait_flags, exchange_no, T.n_regions};
    HeapJobs J = heap_jobs(A, T, n_houses, n_offices, grid_blocks);
    void* args[] = {(void*)&P, (void*)&D, (void*)&A, (void*)&T, (void*)&block_counts, (void*)&ballots, (void*)&nb, (void*)&send, (void*)&stride, (void*)&X,
                    (void*)&recv, (void*)&J, (void*)&W};
    return cudaLaunchCooperativeKernel((const void*)k_travel_exchange, dim3(grid_blocks), dim3(256), args, 0, s);
}

// creation / reset: the level histograms of both heaps and the foreign-slot marks from scratch (tot_* and foreign zeroed by the caller)
void launch_travel_recount(const Params& P, const DevPtrs& D, const TravelPtrs& T, uint32_t n_houses, uint32_t n_offices, cudaStream_t s) {
    k_occ_count_levels<<<blocks_for(n_houses), 256, 0, s>>>(T.occ_house, n_houses, HOUSE_CAP, T.bh_house, T.tot_house);
    k_occ_count_levels<<<blocks_for(n_offices), 256, 0, s>>>(T.occ_office, n_offices, OFFICE_CAP, T.bh_office, T.tot_office);
    k_mark_foreign<<<blocks_for(P.n), 256, 0, s>>>(P, D.st, D.reg, T.foreign);
}

// grid of the cooperative exchange kernels: what is resident at once, capped at `per_sm` blocks per SM
unsigned travel_grid_blocks(int device, int per_sm) {
    int occ_leave = 0, occ_arrive = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_leave, k_travel_leave, 256, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_arrive, k_travel_arrive, 256, 0);
    int occ_x = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_x, k_travel_exchange, 256, 0);
    occ_arrive = std::min(occ_arrive, occ_x);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int occ = std::max(1, std::min(std::min(occ_leave, occ_arrive), per_sm));
    return (unsigned)occ * (unsigned)std::max(1, sms);
}

}  // namespace epi
