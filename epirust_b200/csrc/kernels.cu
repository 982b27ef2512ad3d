// sm_100a kernels of the per-hour agent step.  HBM/L2-bound integer work: no tensor cores.
//
//   k_hospital_scan   first vacant hospital cell in row-major order        (allocation_map.rs:144-147)
//   k_hour<KIND, HOD> one agent per thread: routine, movement proposal against the start-of-hour grid, disease
//                     transition, Counts; atomicMax claim on the target cell (citizen/mod.rs:227-432,
//                     default_disease_handler.rs:31-103, counts.rs:126-140).  KIND: the hour-of-day class
//                     (ROUTINE_START_TIME, ROUTINE_END_TIME, else perform_movements); HOD: the movement hour the kernel is
//                     compiled for (7, 8, 12, 16, 17 or "any other") -- both uniform per launch.
//   k_commit_lanes    two agents per thread, 32 slots apart: lowest-id claimant moves, loser stays; grid bytes updated in place
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

#include "hour_common.cuh"
#include "kernels.h"

namespace epi {


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

// k_hour: agent_hour() of hour_common.cuh, one agent per thread in id order
template <int KIND, bool INJECT, uint32_t HOD = 0>
__global__ void __launch_bounds__(EPI_HBS, KIND == KIND_MOVE ? EPI_MINB * (256 / EPI_HBS) : 4 * (256 / EPI_HBS)) k_hour(Params P, DevPtrs D, uint32_t hour_offset) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    // L2 prefetch of the words the thread one wave ahead loads first, one per 32-byte sector (a population whose arrays stay in
    // the L2 anyway gains nothing from it: 1 M agents run 1 % faster without).  Predicated, not branched, and with the distance
    // as an immediate offset, so the addresses are the ones agent_hour's own loads use.
    const uint32_t pf = P.n >= PREFETCH_MIN_AGENTS && (threadIdx.x & 7u) == 0 && i + PREFETCH_AHEAD < P.n;
    if (KIND == KIND_MOVE)
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %4, 0;\n @p prefetch.global.L2 [%0 + %5];\n @p prefetch.global.L2 [%1 + %5];\n @p prefetch.global.L2 [%2 + %5];\n @p prefetch.global.L2 [%3 + %5];\n}"
                     ::"l"(D.st + i), "l"(D.cell + i), "l"(D.home + i), "l"(D.work + i), "r"(pf), "n"(PREFETCH_AHEAD * 4u));
    else
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p prefetch.global.L2 [%0 + %4];\n @p prefetch.global.L2 [%1 + %4];\n @p prefetch.global.L2 [%2 + %4];\n}"
                     ::"l"(D.st + i), "l"(D.cell + i), "l"(D.home + i), "r"(pf), "n"(PREFETCH_AHEAD * 4u));
    const uint2 clk = load_clock(D.clock);  // hour_base, epoch_base
    GlobalEnv<true> env{P, D, clk.y};
    agent_hour<KIND, INJECT, HOD>(P, D, i, clk.x + hour_offset, env);
}


// k_commit<LAZY = true>: the commit pass of a tile hour (tiles.cu), four consecutive agents per thread.  The proposal and cell words
// arrive as two 128-bit loads, the (up to) four claim words are requested together before anything is stored.
#ifndef EPI_COMMIT_APT
#define EPI_COMMIT_APT 2u  // agents per thread of k_commit_lanes (ms per simulated day at 10 M / 1 M agents: 1: 4.72 / 0.522, 2: 4.54 / 0.511, 4: 4.56 / 0.528, 8: 4.59 / 0.547)
#endif
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
// LAZY (the commit pass of a tile hour, tiles.cu): most proposal words are zero -- the tile kernel settled those agents on chip --
// so the cell words are loaded only where a proposal is pending, a proposal without PROP_MOVE / PROP_DIRTY is a deferred "clear
// this cell" of a tile member that moved in from outside its tile, and consumed proposals are zeroed (the tile kernels write
// only non-zero ones).  zero_props: the next hour is a tile hour, leave prop[] all zero.
template <bool LAZY>
__global__ void __launch_bounds__(256) k_commit(Params P, DevPtrs D, uint32_t hour_offset, uint32_t zero_props) {
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    const uint32_t hour = D.clock->hour_base + hour_offset;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(D.trace, 1, hour);
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
    if (full && LAZY) {
        const uint4 p4 = ld_stream4(D.prop + i0);
        prop[0] = p4.x; prop[1] = p4.y; prop[2] = p4.z; prop[3] = p4.w;
        if ((prop[0] | prop[1] | prop[2] | prop[3]) == 0) return;
        const uint4 c4 = ld_stream4(D.cell + i0);
        c0[0] = c4.x; c0[1] = c4.y; c0[2] = c4.z; c0[3] = c4.w;
    } else if (full) {
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
        if (LAZY && !(prop[j] & (PROP_MOVE | PROP_DIRTY))) {  // deferred clear of the cell a tile member left
            D.grid[P.cell_offset(prop[j] & PROP_CELL_MASK)] = 0;
            continue;
        }
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
    }    if (LAZY || zero_props) {
        if (full) st_stream4(D.prop + i0, 0u, 0u, 0u, 0u);
        else {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (i0 + j < P.n) D.prop[i0 + j] = 0u;
        }
    }
}

// k_commit_lanes: the commit pass of the id-order hours.  The EPI_COMMIT_APT agents of a thread are 32 slots apart (agent = warp base
// + 32 j + lane), so every load / store instruction of a warp covers 32 CONSECUTIVE agents: housemates sit in adjacent lanes and the
// claim words and grid bytes one instruction touches share 128-byte lines (the L1 processes one line per cycle, which is this
// kernel's limit), while a thread still has several claim loads in flight.  Against four consecutive agents per thread (k_commit
// below, 128-bit loads): 96 -> 88 us per launch at 10 M agents, 17.4 -> 14.2 us at 1 M.
__device__ __forceinline__ uint32_t ld_stream1(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(256) k_commit_lanes(Params P, DevPtrs D, uint32_t hour_offset, uint32_t zero_props) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t wbase = (gt >> 5) * (32u * EPI_COMMIT_APT), base = wbase + (gt & 31u);
    const uint32_t hour = D.clock->hour_base + hour_offset;
    if (gt == 0) trace_stamp(D.trace, 1, hour);
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
    if (wbase >= P.n) return;
    uint32_t prop[EPI_COMMIT_APT], c0[EPI_COMMIT_APT];
#pragma unroll
    for (uint32_t j = 0; j < EPI_COMMIT_APT; ++j) {
        const uint32_t idx = base + 32u * j;
        const bool in = idx < P.n;
        prop[j] = in ? ld_stream1(D.prop + idx) : 0u;
        c0[j] = in ? ld_stream1(D.cell + idx) : 0u;
    }
    uint32_t any = 0;
#pragma unroll
    for (uint32_t j = 0; j < EPI_COMMIT_APT; ++j) any |= prop[j];
    if (any == 0) return;
    if (zero_props) {  // the next hour is a tile hour: leave prop[] all zero
#pragma unroll
        for (uint32_t j = 0; j < EPI_COMMIT_APT; ++j)
            if (prop[j]) st_stream(D.prop + base + 32u * j, 0u);
    }
    const uint32_t stamp = hour - D.clock->epoch_base + 1u;
    const uint32_t id_mask = (1u << P.id_bits) - 1u;
    uint32_t cl[EPI_COMMIT_APT];  // every claim word first: one round trip for the four agents
#pragma unroll
    for (uint32_t j = 0; j < EPI_COMMIT_APT; ++j) {
        cl[j] = 0;
        if (prop[j] & PROP_MOVE) cl[j] = __ldcg(D.claim + P.cell_index(prop[j] & PROP_CELL_MASK & CELL_XMASK, (prop[j] & PROP_CELL_MASK) >> CELL_BITS));
    }
#pragma unroll
    for (uint32_t j = 0; j < EPI_COMMIT_APT; ++j) {
        if (prop[j] == 0) continue;
        const uint32_t idx = base + 32u * j;
        const uint32_t byte = (prop[j] >> PROP_BYTE_SHIFT) + 1u;
        const uint32_t at = P.cell_index(c0[j] & CELL_XMASK, c0[j] >> CELL_BITS);
        if (prop[j] & PROP_MOVE) {
            const uint32_t tc = prop[j] & PROP_CELL_MASK;
            // lowest id among the claimants: upcoming.entry(new).or_insert (allocation_map.rs:93-98)
            if (cl[j] == ((stamp << P.id_bits) | (id_mask - idx))) {
                D.grid[at] = 0;
                D.grid[P.cell_index(tc & CELL_XMASK, tc >> CELL_BITS)] = (uint8_t)byte;
                st_stream(D.cell + idx, tc);
                continue;
            }
            // lost: stays at old_cell (allocation_map.rs:99-102); still refresh the byte if it changed
        }
        if (prop[j] & PROP_DIRTY) D.grid[at] = (uint8_t)byte;
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
void launch_commit(const Params& P, const DevPtrs& D, uint32_t hour_offset, bool lazy, bool zero_props, cudaStream_t s) {
    if (lazy) k_commit<true><<<blocks_for((P.n + 3u) / 4u), 256, 0, s>>>(P, D, hour_offset, 1u);
    else k_commit_lanes<<<blocks_for((P.n + 32u * EPI_COMMIT_APT - 1u) / (32u * EPI_COMMIT_APT) * 32u), 256, 0, s>>>(P, D, hour_offset, zero_props ? 1u : 0u);
}
void launch_sleep(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s) { k_sleep<<<blocks_for((P.n + 3u) / 4u), 256, 0, s>>>(P, D, hour_offset); }
void launch_lock(const Params& P, const DevPtrs& D, cudaStream_t s) { k_lock<<<blocks_for(P.n), 256, 0, s>>>(P, D.st); }
void launch_unlock(const Params& P, const DevPtrs& D, cudaStream_t s) { k_unlock<<<blocks_for(P.n), 256, 0, s>>>(P, D.st); }
void launch_vaccinate(const Params& P, const DevPtrs& D, uint64_t thr, uint32_t hour, cudaStream_t s) { k_vaccinate<<<blocks_for(P.n), 256, 0, s>>>(P, D.st, thr, hour); }
void launch_build_grid(const Params& P, const DevPtrs& D, uint32_t* collisions, cudaStream_t s) {
    k_build_grid<<<blocks_for(P.n), 256, 0, s>>>(P, D.cell, D.st, D.grid, collisions);
}

}  // namespace epi
