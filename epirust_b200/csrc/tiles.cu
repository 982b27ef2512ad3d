// Tile kernels of the plain movement hours (h = 9..11, 13..15, 18..22: everybody who can move walks inside current_area).
//
// During these hours an agent's movement and its exposure scan are confined to one small rectangle R (Citizen::move_agent_from
// and on_susceptible clip to current_area: the agent's 10x10 office during the office hours, its 2x2 house in the evening;
// citizen/mod.rs:415-432, default_disease_handler.rs:64-86).  So all agents that can touch the cells of an office (a house) this
// hour are its members, and a CTA that owns a run of x-adjacent offices (houses) and ALL their members can do the whole hour
// on chip:
//   * the tile's grid bytes are staged into shared memory by the TMA unit (one cp.async.bulk per grid row, mbarrier
//     complete_tx), windows are read from there (start-of-hour snapshot);
//   * claims are an atomicMax in shared memory (lowest agent id wins, allocation_map.rs:93-102) -- claim[] in HBM is not touched;
//   * after a __syncthreads the winners are known: cell[] and the tile's grid bytes are updated at once and the tile is written
//     back coalesced -- no proposal word, no k_commit pass for these agents.
// Which agents a CTA processes comes from two device-built orders (cub radix sort by tile, stable so ids ascend inside a tile and
// the per-agent loads stay sector-coalesced): order A = Normal / Essential workers by office tile (office hours), order B =
// everybody but public-transport commuters by house tile (evening hours).  Whoever is not in a tile segment (and every member
// whose rectangle this hour is not its tile's kind: hospital staff, the sick at home, ...) takes the unchanged global path
// (agent_hour with GlobalEnv: windows from global memory, atomicMax on claim[], proposal word for k_commit).
//
// Why this is exact (the tile never misses a claimant):
//   * a cell of office O can only be proposed by an agent whose rectangle is O, i.e. whose work word is O: a member of O's tile
//     when it is a Normal / Essential worker (order A holds all of them); anybody else with that rectangle (never happens in a
//     reference run; crafted states) is in the generic segment, which runs in an EARLIER launch and marks the tile dirty -- a
//     dirty tile falls back to the global path as a whole;
//   * a cell of house H can be proposed by residents of H (members of H's tile, or generic public-transport commuters -> dirty
//     mark) and by agents whose current_area is the whole housing strip (non-working agents between 08:00 and 12:00): a count
//     of those (k_count_housing, taken when the evening phase starts; the plain hours never change current_area) disables the
//     house tiles while it is non-zero;
//   * cells outside R are masked out of every decision (valid_mask), so a tile needs no halo and concurrent write-backs of
//     neighbouring tiles are harmless; effects outside the tile (the old cell of a member that stood elsewhere) are deferred
//     to k_commit through the proposal word.
#include <cub/device/device_radix_sort.cuh>

#include "hour_common.cuh"
#include "kernels.h"

namespace epi {

#ifndef EPI_TILE_THREADS
#define EPI_TILE_THREADS 256  // threads per CTA of the tile kernel: one tile per warp
#endif
#ifndef EPI_TILE_MINB
#define EPI_TILE_MINB 5
#endif
constexpr uint32_t TILE_XPAD = 16;  // bytes in front of / behind a staged row (window loads reach 2 cells out; rows start 16-byte aligned)

__device__ __forceinline__ uint32_t tile_of_origin(const TileGeom& G, int sx, int sy) {
    const uint32_t ux = (uint32_t)(sx - G.ox) / (uint32_t)G.unit, uy = (uint32_t)(sy - G.oy) / (uint32_t)G.unit;
    return uy * (uint32_t)G.chunks + ux / (uint32_t)G.tile_units;
}

// ---- the two orders ---------------------------------------------------------------------------------------------------------
// key of agent slot i: 2 * tile for a member of the tile's local class, 2 * tile + 1 for an agent that rides along with the tile on
// the global path, 2 * n_tiles for the generic segment.
//   order A (cls = RC_OFFICE): Normal / Essential workers by the tile of their office; a non-working agent rides with the first
//     worker of its aligned group of 8 slots (so the 32-byte sectors of the per-agent arrays are fetched once for both);
//     hospital staff and slots without a worker nearby are generic.
//   order B (cls = RC_HOME): everybody by the tile of their house; public-transport commuters (they spend the evening in the
//     transport strip) ride along.
__device__ __forceinline__ bool is_worker(uint32_t s) {
    const uint32_t ws = (s >> ST_WS_SHIFT) & 3u;
    return (s & ST_STATE_MASK) != ST_ABSENT && (ws == WS_NORMAL || ws == WS_ESSENTIAL);
}
__global__ void __launch_bounds__(256) k_tile_keys(Params P, DevPtrs D, TileGeom G, uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t s = D.st[i];
    uint32_t key = 2u * G.n_tiles;
    if ((s & ST_STATE_MASK) != ST_ABSENT) {
        if (G.cls == RC_OFFICE) {
            uint32_t j = i;
            bool found = is_worker(s);
            if (!found && ((s >> ST_WS_SHIFT) & 3u) == WS_NA)
                for (j = i & ~7u; j < min(P.n, (i & ~7u) + 8u); ++j)
                    if (is_worker(D.st[j])) { found = true; break; }
            if (found) {
                const uint32_t w = D.work[j];
                key = 2u * tile_of_origin(G, (int)(w & CELL_XMASK), (int)((w >> CELL_BITS) & CELL_XMASK)) + (j != i);
            }
        } else {
            const uint32_t hm = D.home[i];
            key = 2u * tile_of_origin(G, (int)(hm & CELL_XMASK), (int)((hm >> CELL_BITS) & CELL_XMASK)) + ((is_worker(s) && (s & ST_PT)) ? 1u : 0u);
        }
    }
    keys[i] = min(key, 2u * G.n_tiles);
    ids[i] = i;
}
// start[t] = first position of key >= t in the sorted keys (t = 0 .. n_keys)
__global__ void __launch_bounds__(256) k_tile_starts(const uint32_t* __restrict__ sorted_keys, uint32_t n, uint32_t n_keys, uint32_t* __restrict__ start) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_keys) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < t) lo = mid + 1u; else hi = mid;
    }
    start[t] = lo;
}
// agents whose current_area is the housing strip (they may step into anybody's house)
__global__ void __launch_bounds__(256) k_count_housing(Params P, const uint32_t* __restrict__ st, uint32_t* __restrict__ out) {
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    uint32_t c = 0;
    if (i0 + 3u < P.n) {
        const uint4 v = *reinterpret_cast<const uint4*>(st + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) c += (w[j] & ST_STATE_MASK) != ST_ABSENT && ((w[j] >> ST_AREA_SHIFT) & 7u) == AK_HOUSING;
    } else {
        for (uint32_t j = 0; j < 4u && i0 + j < P.n; ++j) c += (st[i0 + j] & ST_STATE_MASK) != ST_ABSENT && ((st[i0 + j] >> ST_AREA_SHIFT) & 7u) == AK_HOUSING;
    }
    if (__any_sync(0xFFFFFFFFu, c != 0) && c) atomicAdd(out, c);
}

// ---- launch 1: the generic segment (global path), marking the tiles it proposes into ----------------------------------------------
struct MarkEnv : GlobalEnv<false> {
    const TileGeom& G;
    uint32_t* dirty;
    __device__ __forceinline__ MarkEnv(const Params& P_, const DevPtrs& D_, const TileGeom& G_, uint32_t* dirty_) : GlobalEnv<false>{P_, D_, D_.clock->epoch_base}, G(G_), dirty(dirty_) {}
    __device__ __forceinline__ void on_rule(int rcls, const Rect& R, int mode) {
        if (rcls == G.cls && mode != MODE_STAY) dirty[tile_of_origin(G, R.sx, R.sy)] = 1u;
    }
};
__global__ void __launch_bounds__(128, 8) k_hour_list(Params P, DevPtrs D, TileGeom G, TilePtrs TP, uint32_t hour_offset) {
    const uint32_t k = TP.start[2u * G.n_tiles] + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.n) return;
    MarkEnv env(P, D, G, TP.dirty);
    agent_hour<KIND_MOVE, false, 9>(P, D, TP.perm[k], D.clock->hour_base + hour_offset, env);
}

// ---- launch 2: one WARP per tile ----------------------------------------------------------------------------------------------------
// A warp owns a tile: `snap` = its grid bytes at the start of the hour (what every window is read from), `next` = the same bytes
// with this hour's moves applied (written back at the end).  The tile's local members are processed 32 at a time in ascending
// id order, so "lowest id wins" needs no claim words at all: among the 32 lanes of an iteration the lowest lane that proposes a
// cell takes it (__match_any_sync), unless an earlier iteration already did (next[cell] != 0) -- an earlier iteration has
// lower ids, and a cell anybody proposes was vacant at the start of the hour (allocation_map.rs:93-102).
struct TileShared {
    uint8_t *snap, *next;  // [unit + 4 rows][sp]: row r of the tile at (r + 2) * sp, cell x at x - xa + TILE_XPAD
    int x0, y0, w, hgt, xa, sp;
    __device__ __forceinline__ uint32_t at(int x, int y) const { return (uint32_t)((y - y0 + 2) * sp + (x - xa + (int)TILE_XPAD)); }
    __device__ __forceinline__ bool inside(int x, int y) const { return x >= x0 && x < x0 + w && y >= y0 && y < y0 + hgt; }
};
__device__ __forceinline__ Window load_window_shared(const TileShared& T, int cx, int cy) {
    const uint32_t first = T.at(cx - 2, cy - 2);
    const uint32_t sh = (first & 3u) * 8u;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(T.snap + (first & ~3u));
    const uint32_t stride = (uint32_t)T.sp >> 2;
    Window win;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t a = p[k * stride], b = p[k * stride + 1];
        win.l[k] = __funnelshift_r(a, b, sh);
        win.r[k] = __funnelshift_rc(a, b, sh + 8u);
    }
    return win;
}
template <int CLS>
struct TileEnv {
    static constexpr bool stream_loads = false;
    const Params& P;
    const DevPtrs& D;
    const TileShared& T;
    uint32_t* err;
    bool local = false;  // this agent-hour is confined to the tile: windows from snap, settled by the warp
    // what the agent wants, kept until the warp has reconverged
    bool pending = false, moving = false, dirty = false;
    int x = 0, y = 0, tx = 0, ty = 0;
    uint32_t byte = 0;
    __device__ __forceinline__ void on_rule(int rcls, const Rect& R, int) {
        local = rcls == CLS;
        if (local && !(R.sx >= T.x0 && R.ex < T.x0 + T.w && R.sy >= T.y0 && R.ey < T.y0 + T.hgt)) {  // the order is stale: engine bug
            *err = 1u;
            local = false;
        }
    }
    __device__ __forceinline__ Window window(int cx, int cy) const {
        if (local && T.inside(cx, cy)) return load_window_shared(T, cx, cy);
        return load_window(D.grid, P, cx, cy);
    }
    __device__ __forceinline__ void commit(uint32_t i, uint32_t hour, int x_, int y_, int tx_, int ty_, bool dirty_, uint32_t byte_) {
        if (!local) {
            GlobalEnv<false, false> g{P, D, D.clock->epoch_base};
            g.commit(i, hour, x_, y_, tx_, ty_, dirty_, byte_);
            return;
        }
        moving = tx_ != x_ || ty_ != y_;  // the target lies in R, R in the tile
        dirty = dirty_;
        pending = moving || dirty;
        x = x_; y = y_; tx = tx_; ty = ty_; byte = byte_;
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int class_of_kind(uint32_t s) {  // the rectangle class of rect_of(current_area) (hour_common.cuh)
    const uint32_t kind = (s >> ST_AREA_SHIFT) & 7u, ws = (s >> ST_WS_SHIFT) & 3u;
    return kind >= AK_TRANSPORT ? RC_ZONE : (kind == AK_WORK && ws != WS_NA) ? RC_OFFICE : RC_HOME;
}

constexpr int TILE_WARPS = EPI_TILE_THREADS / 32;
template <int CLS>
__global__ void __launch_bounds__(EPI_TILE_THREADS, EPI_TILE_MINB) k_hour_tile(Params P, DevPtrs D, TileGeom G, TilePtrs TP, uint32_t hour_offset) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[TILE_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t tile = blockIdx.x * TILE_WARPS + warp;
    if (tile >= G.n_tiles) return;
    const uint32_t a0 = TP.start[2u * tile], a1 = TP.start[2u * tile + 1u], a2 = TP.start[2u * tile + 2u];
    if (a2 == a0) return;
    const uint32_t hour = D.clock->hour_base + hour_offset;
    // geometry of the tile: units u0 .. u0 + nu - 1 of unit row uy
    const int uy = (int)(tile / (uint32_t)G.chunks), u0 = (int)(tile % (uint32_t)G.chunks) * G.tile_units;
    const int nu = min(G.tile_units, G.units_x - u0);
    TileShared T;
    T.x0 = G.ox + G.unit * u0; T.y0 = G.oy + G.unit * uy; T.w = G.unit * nu; T.hgt = G.unit;
    T.xa = T.x0 & ~15;  // cell_offset(xa, y) is 16-byte aligned: GRID_XOFF and pitch are multiples of 16
    const int row_bytes = (T.x0 + T.w - T.xa + 15) & ~15;
    T.sp = G.sp;
    T.snap = smem + (size_t)warp * (size_t)(2 * (G.unit + 4)) * G.sp;
    T.next = T.snap + (size_t)(G.unit + 4) * G.sp;
    // may the tile be settled on chip this hour?
    uint32_t use = 1u;
    if (lane == 0) {
        use = TP.dirty[tile] == 0u;
        if (CLS == RC_HOME) use = use && *TP.n_housing == 0u;
        if (!use) TP.dirty[tile] = 0u;  // for the next hour
    }
    use = __shfl_sync(0xFFFFFFFFu, use, 0);
    if (CLS == RC_HOME && use) {  // riders (public-transport commuters) whose rectangle is their home after all: everybody on the global path
        bool bad = false;
        for (uint32_t k = a1 + lane; k < a2; k += 32u) {
            const uint32_t s = D.st[TP.perm[k]];
            bad = bad || ((s & ST_STATE_MASK) != ST_ABSENT && class_of_kind(s) == CLS);
        }
        use = !__any_sync(0xFFFFFFFFu, bad);
    }
    if (use && a1 > a0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[warp])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // stage the tile's rows twice (snap and next): TMA bulk copies, completion counted in bytes on the warp's mbarrier
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar[warp])), "r"((uint32_t)(2 * row_bytes * T.hgt)) : "memory");
            for (int r = 0; r < T.hgt; ++r) {
                const uint8_t* src = D.grid + P.cell_offset(T.xa, T.y0 + r);
                const uint32_t off = (uint32_t)(r + 2) * (uint32_t)T.sp + TILE_XPAD;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(T.snap + off)), "l"(src),
                             "r"((uint32_t)row_bytes), "r"(smem_u32(&s_bar[warp]))
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(T.next + off)), "l"(src),
                             "r"((uint32_t)row_bytes), "r"(smem_u32(&s_bar[warp]))
                             : "memory");
            }
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&s_bar[warp])) : "memory");
        TileEnv<CLS> env{P, D, T, TP.misc + 1};
        // two iterations ahead: the slot numbers; one ahead: an L2 prefetch of the slot's words
        uint32_t i_cur = a0 + lane < a1 ? TP.perm[a0 + lane] : 0u;
        uint32_t i_nxt = a0 + 32u + lane < a1 ? TP.perm[a0 + 32u + lane] : 0u;
        for (uint32_t base = a0; base < a1; base += 32u) {
            const uint32_t k = base + lane;
            if (k + 32u < a1) {
                prefetch_l2(D.st + i_nxt);
                prefetch_l2(D.cell + i_nxt);
                prefetch_l2(CLS == RC_OFFICE ? D.work + i_nxt : D.home + i_nxt);
            }
            const uint32_t i_nn = k + 64u < a1 ? TP.perm[k + 64u] : 0u;
            const uint32_t i = i_cur;
            i_cur = i_nxt;
            i_nxt = i_nn;
            env.local = false;
            env.pending = false;
            if (k < a1) agent_hour<KIND_MOVE, false, 9>(P, D, i, hour, env);
            __syncwarp();
            // lowest id among the claimants: upcoming.entry(new).or_insert (allocation_map.rs:93-98)
            const bool claims = env.pending && env.moving;
            const uint32_t t_at = claims ? T.at(env.tx, env.ty) : 0u;
            const uint32_t group = __match_any_sync(0xFFFFFFFFu, claims ? t_at : (0x80000000u | lane));
            const bool won = claims && (uint32_t)(__ffs(group) - 1) == lane && T.next[t_at] == 0;
            __syncwarp();
            if (env.pending) {
                const bool in_old = T.inside(env.x, env.y);
                if (won) {
                    D.cell[i] = ((uint32_t)env.ty << CELL_BITS) | (uint32_t)env.tx;
                    T.next[t_at] = (uint8_t)env.byte;
                    if (in_old) T.next[T.at(env.x, env.y)] = 0;
                    else D.prop[i] = (1u << PROP_BYTE_SHIFT) | ((uint32_t)env.y << CELL_BITS) | (uint32_t)env.x;  // somebody else's cell: k_commit clears it after this kernel
                } else if (env.dirty) {  // stays at old_cell (allocation_map.rs:99-102) with a new grid byte
                    if (in_old) T.next[T.at(env.x, env.y)] = (uint8_t)env.byte;
                    else D.prop[i] = PROP_DIRTY | ((env.byte - 1u) << PROP_BYTE_SHIFT);
                }
            }
            __syncwarp();
        }
    }
    // the riders (and the local members of a tile that is not settled on chip this hour): global path
    {
        GlobalEnv<false, false> genv{P, D, D.clock->epoch_base};
        const uint32_t g0 = use ? a1 : a0;
        uint32_t i_cur = g0 + lane < a2 ? TP.perm[g0 + lane] : 0u;
        uint32_t i_nxt = g0 + 32u + lane < a2 ? TP.perm[g0 + 32u + lane] : 0u;
        for (uint32_t k = g0 + lane; k < a2; k += 32u) {
            if (k + 32u < a2) {
                prefetch_l2(D.st + i_nxt);
                prefetch_l2(D.cell + i_nxt);
                prefetch_l2(D.home + i_nxt);
            }
            const uint32_t i_nn = k + 64u < a2 ? TP.perm[k + 64u] : 0u;
            const uint32_t i = i_cur;
            i_cur = i_nxt;
            i_nxt = i_nn;
            agent_hour<KIND_MOVE, false, 9>(P, D, i, hour, genv);
        }
    }
    if (use && a1 > a0) {
        __syncwarp();
        // write the tile back (its own cells only: the alignment padding belongs to the neighbours)
        for (int r = 0; r < T.hgt; ++r) {
            uint8_t* dst = D.grid + P.cell_offset(T.x0, T.y0 + r);
            const uint8_t* srow = T.next + T.at(T.x0, T.y0 + r);
            for (int c = (int)lane; c < T.w; c += 32) dst[c] = srow[c];
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
size_t tile_shared_bytes(const TileGeom& G) { return (size_t)TILE_WARPS * 2 * (size_t)(G.unit + 4) * G.sp; }

size_t tile_sort_temp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 32);
    return bytes;
}

// keys_a / keys_b / ids: scratch of n words each; temp: tile_sort_temp_bytes(n).  Fills TP.perm and TP.start (n_tiles + 2 words).
cudaError_t build_tile_order(const Params& P, const DevPtrs& D, const TileGeom& G, const TilePtrs& TP, uint32_t* keys_a, uint32_t* keys_b, uint32_t* ids, void* temp,
                             size_t temp_bytes, cudaStream_t s) {
    const unsigned blocks = (P.n + 255u) / 256u;
    k_tile_keys<<<blocks, 256, 0, s>>>(P, D, G, keys_a, ids);
    int bits = 1;
    while ((1u << bits) <= 2u * G.n_tiles) ++bits;
    cudaError_t r = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, ids, TP.perm, (int)P.n, 0, bits, s);
    if (r != cudaSuccess) return r;
    k_tile_starts<<<(2u * G.n_tiles + 2u + 255u) / 256u, 256, 0, s>>>(keys_b, P.n, 2u * G.n_tiles + 1u, TP.start);
    return cudaGetLastError();
}

void launch_count_housing(const Params& P, const DevPtrs& D, uint32_t* out, cudaStream_t s) {
    cudaMemsetAsync(out, 0, sizeof(uint32_t), s);
    k_count_housing<<<((P.n + 3u) / 4u + 255u) / 256u, 256, 0, s>>>(P, D.st, out);
}

// one plain hour: the generic segment (n_generic_bound >= its length), then the tiles
unsigned launch_hour_tiles(const Params& P, const DevPtrs& D, const TileGeom& G, const TilePtrs& TP, uint32_t n_generic_bound, uint32_t hour_offset, cudaStream_t s) {
    unsigned launches = 0;
    if (n_generic_bound) {
        k_hour_list<<<(n_generic_bound + 127u) / 128u, 128, 0, s>>>(P, D, G, TP, hour_offset);
        ++launches;
    }
    const size_t shared = tile_shared_bytes(G);
    const unsigned blocks = (G.n_tiles + TILE_WARPS - 1u) / TILE_WARPS;
    if (G.cls == RC_OFFICE) k_hour_tile<RC_OFFICE><<<blocks, EPI_TILE_THREADS, shared, s>>>(P, D, G, TP, hour_offset);
    else k_hour_tile<RC_HOME><<<blocks, EPI_TILE_THREADS, shared, s>>>(P, D, G, TP, hour_offset);
    return launches + 1;
}

cudaError_t tiles_configure(const TileGeom& office, const TileGeom& house) {
    cudaError_t r = cudaFuncSetAttribute(k_hour_tile<RC_OFFICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_shared_bytes(office));
    if (r != cudaSuccess) return r;
    return cudaFuncSetAttribute(k_hour_tile<RC_HOME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_shared_bytes(house));
}

}  // namespace epi
