// sm_100a kernels of the daily traveller exchange (north_star: "packs and unpacks migrating agents on device").
// Replaces CitizenLocationMap::simulate's traveller selection (allocation_map.rs:104-120), remove_migrators /
// remove_commuters (:165-212), assimilate_migrators / assimilate_commuters (:214-277), the allotment of migrators to regions
// (engine_migration_plan.rs:51-77, migrators_by_engine.rs:34-55), commuters_by_region.rs:59-78 and the house / office
// occupancy heaps (grid.rs:47-80, 279-341).  Everything the reference does sequentially on the host is done here on the
// device with the same result; the host only learns the per-region record counts (it needs them for the collective).
//
// Leaving (epi_travel_pack):
//   k_travel_flag_count / k_travel_scan / k_travel_scatter   ordered compaction of the leaving candidates (ascending slot)
//   k_travel_plan      records per destination: migrators by position in the list (floor(share * total) from the front,
//                      the surplus stays), commuters by their work / home region; segment headers
//   k_travel_rank      commuters: stable index inside the destination's segment
//   k_travel_pack      record -> segment of the destination; the agent leaves: cell vacated, slot pushed on the free stack,
//                      Counts decremented, its house / office lose an occupant (migrators)
// Arriving (epi_travel_unpack):
//   k_travel_gather    segments (one per source region) -> contiguous arrival list, ordered by source region
//   k_travel_wscan     rank of every arriving migrator among the working ones (they also take an office)
//   k_occ_*            "water filling": the pop sequence of the reference's BinaryHeap of areas for K arrivals, in parallel.
//                      The heap pops the least occupied area, ties to the greatest (start.x, start.y); a popped area
//                      returns with one more occupant.  So the pops run level by level: first every area of the lowest
//                      occupancy L in tie order, then every area whose occupancy was <= L + 1 in tie order, and so on.
//                      occ[] is stored in tie order, so "the q-th area with occupancy <= l" is a prefix count.
//   k_travel_install   arrival k -> the k-th slot from the top of the free stack (Citizen::from_migrator / from_commuter)
//   k_travel_propose / k_travel_grant   placement rounds: distinct vacant cells of the arrival strip, lowest arrival wins
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
// Ordered compaction of the leaving candidates.  Pass 1 evaluates the flag of every slot once, keeps the warp ballots and
// counts per block of 1024 slots; one block scans the counts; pass 2 places the flagged slots from the ballots (it reads
// reg[] again only for the few flagged agents).
constexpr uint32_t SEL_BLOCK = 1024;  // slots per count / offset entry
// 256 threads x 4 consecutive slots (two 128-bit loads per thread); eight threads assemble one 32-slot ballot word
__global__ void __launch_bounds__(256) k_travel_flag_count(Params P, DevPtrs D, TravelArgs A, uint32_t* __restrict__ block_counts, uint32_t* __restrict__ ballots) {
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
    uint32_t nib = 0;
    if (i0 + 3u < P.n) {
        const uint4 s4 = __ldcs(reinterpret_cast<const uint4*>(D.st + i0)), r4 = __ldcs(reinterpret_cast<const uint4*>(D.reg + i0));
        nib = (travel_flag(P, A, i0, s4.x, r4.x) ? 1u : 0u) | (travel_flag(P, A, i0 + 1u, s4.y, r4.y) ? 2u : 0u) |
              (travel_flag(P, A, i0 + 2u, s4.z, r4.z) ? 4u : 0u) | (travel_flag(P, A, i0 + 3u, s4.w, r4.w) ? 8u : 0u);
    } else {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (i0 + j < P.n && travel_flag(P, A, i0 + j, D.st[i0 + j], D.reg[i0 + j])) nib |= 1u << j;
    }
    uint32_t w = nib << ((threadIdx.x & 7u) * 4u);
    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
    if ((threadIdx.x & 7u) == 0) ballots[i0 >> 5] = w;  // written for every group of the launch, also beyond P.n
    uint32_t c = (threadIdx.x & 7u) == 0 ? (uint32_t)__popc(w) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(&s_count, c);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = s_count;
}

// exclusive scan of the block counts by ONE block (launched 3 times per simulated day)
__global__ void __launch_bounds__(1024) k_travel_scan(uint32_t* __restrict__ block_counts, uint32_t n_blocks, TravelVars* __restrict__ tv) {
    __shared__ uint32_t warp_sums[32];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < n_blocks ? block_counts[idx] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_sums, &total);
        if (idx < n_blocks) block_counts[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        tv->total = carry;
        tv->n_send = 0; tv->n_working = 0; tv->pending = 0; tv->abort = 0;
    }
}

// one warp per block of 1024 slots: lane l owns ballot word l of the block
__global__ void __launch_bounds__(256) k_travel_scatter(Params P, DevPtrs D, TravelArgs A, TravelPtrs T, const uint32_t* __restrict__ block_offsets,
                                                         const uint32_t* __restrict__ ballots, uint32_t n_blocks) {
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (g >= n_blocks) return;
    uint32_t b = ballots[g * 32u + lane];
    uint32_t incl = (uint32_t)__popc(b);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((int)lane >= o) incl += y;
    }
    uint32_t at = block_offsets[g] + incl - (uint32_t)__popc(b);
    while (b) {
        const uint32_t i = (g * 32u + lane) * 32u + (uint32_t)__ffs((int)b) - 1u;
        b &= b - 1u;
        if (at < T.list_cap) {
            T.list_slot[at] = i;
            uint32_t dest = 0;
            if (A.kind == TRAVEL_COMMUTE) {  // Citizen::is_commuter: the work region in the morning, the home region in the evening
                const uint32_t reg = D.reg[i];
                dest = A.hour_of_day == 7u ? (reg >> 8) & 0xFFu : reg & 0xFFu;
            }
            T.list_dest[at] = dest;
        } else atomicOr(&T.tv->err, TERR_LIST_OVERFLOW);
        ++at;
    }
}

// Records per destination and the segment headers.  Migrators: EngineMigrationPlan::alloc_outgoing_to_regions
// (engine_migration_plan.rs:51-77) + MigratorsByRegion::alloc_citizens (migrators_by_engine.rs:34-55): regions in plan
// order take floor(share * total) from the front of the list, the rest stays.  Commuters:
// CommutersByRegion::get_commuters_by_region (commuters_by_region.rs:59-78).
__global__ void __launch_bounds__(1024) k_travel_plan(Params P, TravelArgs A, TravelPtrs T, TravelRecord* __restrict__ send, uint32_t stride) {
    __shared__ uint32_t hist[TRAVEL_MAX_REGIONS];
    TravelVars* tv = T.tv;
    const uint32_t R = (uint32_t)T.n_regions;
    const uint32_t total = min(tv->total, T.list_cap);
    for (uint32_t r = threadIdx.x; r < TRAVEL_MAX_REGIONS; r += blockDim.x) hist[r] = 0;
    __syncthreads();
    if (A.kind == TRAVEL_COMMUTE) {
        for (uint32_t p = threadIdx.x; p < total; p += blockDim.x) {
            const uint32_t d = T.list_dest[p];
            if (d < R) atomicAdd(&hist[d], 1u);
            else atomicOr(&tv->err, TERR_BAD_REGION);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
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
            for (uint32_t to = 0; to < R; ++to) { tv->cnt[to] = hist[to]; tv->base[to] = front; front += hist[to]; }
        }
        tv->n_send = front;
    }
    __syncthreads();
    for (uint32_t to = threadIdx.x; to < R; to += blockDim.x) {
        if (tv->cnt[to] + 1u > (T.seg_cap ? min(T.seg_cap[to], stride) : stride)) atomicOr(&tv->err, TERR_SEGMENT_OVERFLOW);
        TravelRecord h{};
        h.st = min(tv->cnt[to], stride - 1u);  // header: records in this segment
        h.from = (uint32_t)P.region;
        send[(size_t)to * stride] = h;
    }
    // A list or segment overflow is known by now: k_travel_pack then removes nobody (the region stays intact, the host reports
    // the error), and the headers promise no records.
    __syncthreads();
    if (threadIdx.x == 0) tv->abort = tv->err;
    __syncthreads();
    if (tv->err)
        for (uint32_t to = threadIdx.x; to < R; to += blockDim.x) send[(size_t)to * stride].st = 0;
}

// commuters: list_pos[p] = number of earlier candidates with the same destination (block d serves destination d)
__global__ void __launch_bounds__(1024) k_travel_rank(TravelPtrs T) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t d = blockIdx.x, total = min(T.tv->total, T.list_cap);
    uint32_t carry = 0;
    for (uint32_t base = 0; base < total; base += 1024u) {
        const uint32_t p = base + threadIdx.x;
        const uint32_t f = (p < total && T.list_dest[p] == d) ? 1u : 0u;
        uint32_t sum;
        const uint32_t ex = block_exclusive_scan(f, warp_sums, &sum);
        if (f) T.list_pos[p] = carry + ex;
        carry += sum;
    }
}

// the agent leaves: record written, cell vacated, slot emptied and pushed, Counts decremented (decrement_counts,
// allocation_map.rs:291-301), occupancies of its house / office decremented (remove_migrators, :165-192)
__global__ void __launch_bounds__(256) k_travel_pack(Params P, DevPtrs D, TravelArgs A, TravelPtrs T, TravelRecord* __restrict__ send, uint32_t stride) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    TravelVars* tv = T.tv;
    const uint32_t free_top = tv->free_top;
    const uint32_t total = min(tv->total, T.list_cap);
    if (p >= total || tv->abort) return;
    uint32_t dest, j;
    if (A.kind == TRAVEL_MIGRATE) {
        if (p >= tv->n_send) return;  // surplus candidates stay
        dest = 0;
        while (!(p >= tv->base[dest] && p < tv->base[dest] + tv->cnt[dest])) ++dest;
        j = p - tv->base[dest];
    } else {
        dest = T.list_dest[p];
        if (dest >= (uint32_t)T.n_regions) return;
        j = T.list_pos[p];
    }
    const uint32_t i = T.list_slot[p];
    const uint32_t s = D.st[i];
    TravelRecord r;
    r.st = s; r.t0 = D.t0[i]; r.home = D.home[i]; r.work = D.work[i]; r.reg = D.reg[i]; r.slot = i; r.from = (uint32_t)P.region; r.pad = 0;
    if (j + 1u < stride) send[(size_t)dest * stride + 1u + j] = r;
    D.grid[P.cell_offset(D.cell[i])] = 0;
    D.st[i] = ST_ABSENT;
    D.prop[i] = 0;
    atomicSub(D.tot + count_category(s), 1u);
    // free_slots.push_back: migrators in send order (remove_migrators walks the per-region lists), commuters in selection order
    T.free_stack[free_top + p] = i;
    if (A.kind == TRAVEL_MIGRATE) {
        const uint32_t old = atomicSub(&T.occ_house[house_rank(P, r.home)], 1u);
        if (old == OCC_ABSENT || old == 0u) atomicOr(&tv->err, TERR_NO_HOUSE);  // "Could not find house"
        if (((s >> ST_WS_SHIFT) & 3u) != WS_NA) {
            const uint32_t oldo = atomicSub(&T.occ_office[office_rank(P, r.work)], 1u);
            if (oldo == OCC_ABSENT || oldo == 0u) atomicOr(&tv->err, TERR_NO_OFFICE);
        }
    }
}

// the free-slot stack grows by the leavers (after k_travel_pack) / shrinks by the arrivals (after the last placement round)
__global__ void k_travel_stack_moved(TravelVars* tv, int arrivals) {
    if (arrivals) tv->free_top -= tv->n_in;
    else if (!tv->abort) tv->free_top += tv->n_send;
}

// ---- arriving -----------------------------------------------------------------------------------------------------------
// recv: one segment per source region (header + records) -> arrivals[k], k ascending in (source region, index)
// n_in = the sum of the segment headers (tv->cnt[s] keeps the header of source s for the host)
__global__ void __launch_bounds__(256) k_travel_count_in(TravelPtrs T, const TravelRecord* __restrict__ recv, uint32_t stride) {
    if (threadIdx.x != 0) return;
    const uint32_t free_top = T.tv->free_top;
    uint32_t n = 0;
    for (int s = 0; s < T.n_regions; ++s) {
        const uint32_t c = min(recv[(size_t)s * stride].st, stride - 1u);
        T.tv->cnt[s] = c;
        n += c;
    }
    if (n > T.list_cap) { atomicOr(&T.tv->err, TERR_LIST_OVERFLOW); n = 0; }
    if (n > free_top) { atomicOr(&T.tv->err, TERR_NO_SLOTS); n = 0; }  // "region is out of agent slots"
    T.tv->n_in = n;
    T.tv->n_working = 0;
    T.tv->pending = 0;
}
__global__ void __launch_bounds__(256) k_travel_gather(TravelPtrs T, const TravelRecord* __restrict__ recv, uint32_t stride) {
    const uint32_t s = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (T.tv->n_in == 0) return;
    if (j >= T.tv->cnt[s]) return;
    uint32_t k = j;
    for (uint32_t q = 0; q < s; ++q) k += T.tv->cnt[q];
    T.arrivals[k] = recv[(size_t)s * stride + 1u + j];
    T.placed[k] = 0;
}

// migrators: arr_widx[k] = number of working arrivals before k; tv->n_working = their total (they pop an office each)
__global__ void __launch_bounds__(1024) k_travel_wscan(TravelPtrs T) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t n_in = T.tv->n_in;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_in; base += 1024u) {
        const uint32_t k = base + threadIdx.x;
        const uint32_t f = (k < n_in && ((T.arrivals[k].st >> ST_WS_SHIFT) & 3u) != WS_NA) ? 1u : 0u;
        uint32_t sum;
        const uint32_t ex = block_exclusive_scan(f, warp_sums, &sum);
        if (k < n_in) T.arr_widx[k] = f ? carry + ex : 0xFFFFFFFFu;
        carry += sum;
    }
    if (threadIdx.x == 0) T.tv->n_working = carry;
}

// ---- water filling -------------------------------------------------------------------------------------------------------
// per-block histogram of the occupancy levels 0 .. CAP-1 (areas at CAP or more, or absent, cannot be popped usefully)
template <uint32_t CAP>
__global__ void __launch_bounds__(256) k_occ_block_hist(const uint32_t* __restrict__ occ, uint32_t n, uint32_t* __restrict__ bh) {
    __shared__ uint32_t h[CAP];
    for (uint32_t l = threadIdx.x; l < CAP; l += blockDim.x) h[l] = 0;
    __syncthreads();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t o = r < n ? occ[r] : OCC_ABSENT;
    if (CAP <= 8) {  // few levels: block-wide counts instead of shared atomics on a handful of addresses
#pragma unroll
        for (uint32_t l = 0; l < CAP; ++l) {
            const int c = __syncthreads_count(o == l);
            if (threadIdx.x == 0) h[l] = (uint32_t)c;
        }
    } else if (o < CAP) atomicAdd(&h[o], 1u);
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < CAP; l += blockDim.x) bh[(size_t)blockIdx.x * CAP + l] = h[l];
}
// the level structure of K pops.  k_src: K is read from *k_src when non-null (a device-side count), else k_arg.
template <uint32_t CAP>
__global__ void __launch_bounds__(1024) k_occ_plan(const uint32_t* __restrict__ bh, uint32_t n_blocks, FillPlan* __restrict__ plan, const uint32_t* k_src,
                                                    TravelVars* tv, uint32_t err_bit) {
    __shared__ uint32_t H[CAP];
    for (uint32_t l = threadIdx.x; l < CAP; l += blockDim.x) H[l] = 0;
    __syncthreads();
    {   // thread t sums level t % CAP over the blocks t / CAP, t / CAP + groups, ...
        const uint32_t groups = 1024u / CAP;
        if (threadIdx.x < groups * CAP) {
            const uint32_t l = threadIdx.x % CAP;
            uint32_t sum = 0;
            for (uint32_t b = threadIdx.x / CAP; b < n_blocks; b += groups) sum += bh[(size_t)b * CAP + l];
            if (sum) atomicAdd(&H[l], sum);
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const uint32_t K = *k_src;
    plan->K = K; plan->L = 0; plan->l_last = 0; plan->last_count = 0;
    if (K == 0) return;
    uint32_t L = 0;
    while (L < CAP && H[L] == 0) ++L;
    uint64_t cum = 0, n_le = 0;
    bool done = false;
    for (uint32_t l = L; l < CAP; ++l) {
        n_le += H[l];
        plan->start[l] = (uint32_t)cum;
        if (cum + n_le >= K) { plan->l_last = l; plan->last_count = K - (uint32_t)cum; done = true; break; }
        cum += n_le;
    }
    plan->L = L;
    if (!done) {  // "Couldn't find any house / offices with free space!" (allocation_map.rs:222, :254)
        plan->K = 0;
        atomicOr(&tv->err, err_bit);
    }
}
// pref[l * n_blocks + b] = number of areas with occupancy <= l in blocks before b, for the levels the plan uses
template <uint32_t CAP>
__global__ void __launch_bounds__(1024) k_occ_prefix(const uint32_t* __restrict__ bh, uint32_t n_blocks, const FillPlan* __restrict__ plan, uint32_t* __restrict__ pref) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t l = blockIdx.x;
    if (plan->K == 0 || l < plan->L || l > plan->l_last) return;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t b = base + threadIdx.x;
        uint32_t v = 0;
        if (b < n_blocks)
            for (uint32_t j = plan->L; j <= l; ++j) v += bh[(size_t)b * CAP + j];
        uint32_t sum;
        const uint32_t ex = block_exclusive_scan(v, warp_sums, &sum);
        if (b < n_blocks) pref[(size_t)l * n_blocks + b] = carry + ex;
        carry += sum;
    }
}
// pop p -> the area it returns (index in tie order).  One warp per arrival.
template <uint32_t CAP>
__global__ void __launch_bounds__(256) k_occ_assign(const uint32_t* __restrict__ occ, uint32_t n, uint32_t n_blocks, const FillPlan* __restrict__ plan,
                                                     const uint32_t* __restrict__ pref, uint32_t* __restrict__ out, const uint32_t* __restrict__ pop_index,
                                                     const uint32_t* __restrict__ n_arrivals) {
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (k >= *n_arrivals) return;
    const uint32_t p = pop_index ? pop_index[k] : k;  // which pop of this heap serves arrival k (0xFFFFFFFF: none)
    if (p >= plan->K) {
        if (lane == 0) out[k] = 0xFFFFFFFFu;
        return;
    }
    uint32_t l = plan->L;
    while (l < plan->l_last && p >= plan->start[l + 1]) ++l;
    const uint32_t q = p - plan->start[l];
    const uint32_t* pl = pref + (size_t)l * n_blocks;
    uint32_t lo = 0, hi = n_blocks;  // largest b with pl[b] <= q
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pl[mid] <= q) lo = mid; else hi = mid;
    }
    uint32_t left = q - pl[lo];  // the left-th area with occupancy <= l inside block lo (256 areas)
    uint32_t found = 0xFFFFFFFFu;
    for (uint32_t c = 0; c < 8; ++c) {
        const uint32_t r = lo * 256u + c * 32u + lane;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, r < n && occ[r] <= l);
        const uint32_t cnt = (uint32_t)__popc(m);
        if (left < cnt) {
            // position of the left-th set bit of m
            unsigned mm = m;
            for (uint32_t t = 0; t < left; ++t) mm &= mm - 1u;
            found = lo * 256u + c * 32u + (uint32_t)(__ffs(mm) - 1);
            break;
        }
        left -= cnt;
    }
    if (lane == 0) out[k] = found;
}
// every area that was popped comes back with one more occupant per pop
template <uint32_t CAP>
__global__ void __launch_bounds__(256) k_occ_update(uint32_t* __restrict__ occ, uint32_t n, uint32_t n_blocks, const FillPlan* __restrict__ plan, const uint32_t* __restrict__ pref) {
    __shared__ uint32_t warp_sums[32];
    if (plan->K == 0) return;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t o = r < n ? occ[r] : OCC_ABSENT;
    const uint32_t L = plan->L, l_last = plan->l_last;
    const uint32_t f = o <= l_last ? 1u : 0u;
    uint32_t sum;
    const uint32_t ex = block_exclusive_scan(f, warp_sums, &sum);
    if (!f) return;
    const uint32_t rank_last = pref[(size_t)l_last * n_blocks + blockIdx.x] + ex;
    occ[r] = o + (l_last - max(o, L)) + (rank_last < plan->last_count ? 1u : 0u);
}

// arrival k -> the k-th slot from the top of the free stack.  Citizen::from_migrator / from_commuter (citizen/mod.rs:113-154):
// immunity, vaccinated, uses_public_transport and the disease state travel; hospitalized, isolated, work_quarantined reset;
// work status becomes NA (migrator) or Normal (commuter); current_area = the housing strip.  The cell is assigned by the
// placement rounds.
__global__ void __launch_bounds__(256) k_travel_install(Params P, DevPtrs D, TravelArgs A, TravelPtrs T) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T.tv->n_in) return;
    const uint32_t free_top = T.tv->free_top;
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
    atomicAdd(D.tot + count_category(s), 1u);
}

// ---- placement rounds (select_starting_points, allocation_map.rs:339-347) ------------------------------------------
// In round a every still-unplaced arrival k walks its own candidate sequence Philox(seed, k, hour, DOM_ARRIVAL) blocks
// a * PLACE_TRIES .. a * PLACE_TRIES + PLACE_TRIES - 1 (x in [sx, ex), y in [sy, ey)) and proposes the FIRST candidate that is
// vacant now (start-of-hour occupants and earlier rounds' winners count as occupied); of several proposals for one cell the
// lowest k wins; a loser, or an arrival whose candidates were all occupied, tries again in the next round.
constexpr uint32_t PLACE_TRIES = 8;
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

__global__ void __launch_bounds__(256) k_travel_round_begin(TravelPtrs T) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx <= T.table_mask) { T.table_keys[idx] = 0u; T.table_vals[idx] = 0xFFFFFFFFu; }
    if (idx == 0) T.tv->pending = 0;
}
__global__ void __launch_bounds__(256) k_travel_propose(Params P, DevPtrs D, TravelArgs A, TravelPtrs T, uint32_t attempt) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T.tv->n_in || T.placed[k]) return;
    const uint32_t c = first_vacant_candidate(P, D, A, k, attempt);
    T.list_pos[k] = c;  // this round's proposal (list_pos is free during unpack)
    if (c == 0xFFFFFFFFu) return;
    uint32_t slot = hash_cell(c) & T.table_mask;
    for (;;) {
        const uint32_t old = atomicCAS(&T.table_keys[slot], 0u, c + 1u);
        if (old == 0u || old == c + 1u) { atomicMin(&T.table_vals[slot], k); return; }
        slot = (slot + 1u) & T.table_mask;
    }
}
__global__ void __launch_bounds__(256) k_travel_grant(Params P, DevPtrs D, TravelArgs A, TravelPtrs T) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T.tv->n_in || T.placed[k]) return;
    const uint32_t free_top = T.tv->free_top;
    const uint32_t c = T.list_pos[k];
    bool won = false;
    if (c != 0xFFFFFFFFu) {
        uint32_t slot = hash_cell(c) & T.table_mask;
        for (;;) {
            const uint32_t key = T.table_keys[slot];
            if (key == 0u) break;
            if (key == c + 1u) { won = T.table_vals[slot] == k; break; }
            slot = (slot + 1u) & T.table_mask;
        }
    }
    if (won) {
        const uint32_t i = T.free_stack[free_top - 1u - k];
        D.cell[i] = c;
        D.grid[P.cell_offset(c)] = (uint8_t)cell_byte(P, D.st[i]);
        T.placed[k] = 1;
    } else {
        atomicAdd(&T.tv->pending, 1u);
    }
}

// ---- launchers ---------------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint32_t n) { return (n + 255u) / 256u; }

unsigned launch_travel_leave(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, uint32_t* block_counts, TravelRecord* send,
                             uint32_t stride, cudaStream_t s) {
    const unsigned nb = (P.n + SEL_BLOCK - 1u) / SEL_BLOCK;
    uint32_t* ballots = block_counts + nb + 1;  // the host allocates both in one array
    k_travel_flag_count<<<nb, 256, 0, s>>>(P, D, A, block_counts, ballots);
    k_travel_scan<<<1, 1024, 0, s>>>(block_counts, nb, T.tv);
    k_travel_scatter<<<(nb + 7u) / 8u, 256, 0, s>>>(P, D, A, T, block_counts, ballots, nb);
    k_travel_plan<<<1, 1024, 0, s>>>(P, A, T, send, stride);
    unsigned launches = 6;
    if (A.kind == TRAVEL_COMMUTE) { k_travel_rank<<<(unsigned)T.n_regions, 1024, 0, s>>>(T); ++launches; }
    k_travel_pack<<<blocks_for(T.list_cap), 256, 0, s>>>(P, D, A, T, send, stride);
    k_travel_stack_moved<<<1, 1, 0, s>>>(T.tv, 0);
    return launches;
}

template <uint32_t CAP>
static unsigned water_fill(uint32_t* occ, uint32_t n, uint32_t* bh, uint32_t* pref, FillPlan* plan, const uint32_t* k_src, TravelVars* tv, uint32_t err_bit, uint32_t* out,
                           const uint32_t* pop_index, uint32_t max_arrivals, cudaStream_t s) {
    const unsigned nb = blocks_for(n);
    k_occ_block_hist<CAP><<<nb, 256, 0, s>>>(occ, n, bh);
    k_occ_plan<CAP><<<1, 1024, 0, s>>>(bh, nb, plan, k_src, tv, err_bit);
    k_occ_prefix<CAP><<<CAP, 1024, 0, s>>>(bh, nb, plan, pref);
    k_occ_assign<CAP><<<blocks_for(max_arrivals * 32u), 256, 0, s>>>(occ, n, nb, plan, pref, out, pop_index, &tv->n_in);
    k_occ_update<CAP><<<nb, 256, 0, s>>>(occ, n, nb, plan, pref);
    return 5;
}

// max_arrivals: upper bound of the arrivals (grid size); the kernels read the actual number from the segment headers
unsigned launch_travel_arrive(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, const TravelRecord* recv, uint32_t stride,
                              uint32_t max_arrivals, uint32_t n_houses, uint32_t n_offices, cudaStream_t s) {
    unsigned launches = 2;
    k_travel_count_in<<<1, 32, 0, s>>>(T, recv, stride);
    k_travel_gather<<<dim3(blocks_for(stride), (unsigned)T.n_regions), 256, 0, s>>>(T, recv, stride);
    if (A.kind == TRAVEL_MIGRATE) {
        // assimilate_migrators (allocation_map.rs:214-243): every arrival pops a house, the working ones an office as well
        k_travel_wscan<<<1, 1024, 0, s>>>(T);
        ++launches;
        launches += water_fill<HOUSE_CAP>(T.occ_house, n_houses, T.bh_house, T.pref_house, T.plan_house, &T.tv->n_in, T.tv, TERR_HOUSES_FULL, T.arr_house, nullptr, max_arrivals, s);
        launches += water_fill<OFFICE_CAP>(T.occ_office, n_offices, T.bh_office, T.pref_office, T.plan_office, &T.tv->n_working, T.tv, TERR_OFFICES_FULL, T.arr_office, T.arr_widx,
                                           max_arrivals, s);
    } else if (A.hour == 7u) {
        // assimilate_commuters (allocation_map.rs:245-277): an office is assigned at the absolute hour 7 only (:260)
        launches += water_fill<OFFICE_CAP>(T.occ_office, n_offices, T.bh_office, T.pref_office, T.plan_office, &T.tv->n_in, T.tv, TERR_OFFICES_FULL, T.arr_office, nullptr,
                                           max_arrivals, s);
    }
    k_travel_install<<<blocks_for(max_arrivals), 256, 0, s>>>(P, D, A, T);
    return launches + 1;
}

unsigned launch_travel_rounds(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, uint32_t max_arrivals, uint32_t first_attempt, uint32_t n_rounds,
                              cudaStream_t s) {
    for (uint32_t a = 0; a < n_rounds; ++a) {
        k_travel_round_begin<<<blocks_for(T.table_mask + 1u), 256, 0, s>>>(T);
        k_travel_propose<<<blocks_for(max_arrivals), 256, 0, s>>>(P, D, A, T, first_attempt + a);
        k_travel_grant<<<blocks_for(max_arrivals), 256, 0, s>>>(P, D, A, T);
    }
    return 3 * n_rounds;
}
void launch_travel_arrivals_done(const TravelPtrs& T, cudaStream_t s) { k_travel_stack_moved<<<1, 1, 0, s>>>(T.tv, 1); }

}  // namespace epi
