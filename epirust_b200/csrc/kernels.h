// Launchers of the sm_100a kernels in kernels.cu (callable from plain C++ translation units).
#pragma once
#include <cuda_runtime.h>

#include "layout.h"

namespace epi {
void launch_hospital_scan(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_hour(const Params& P, const DevPtrs& D, uint32_t hour_of_day, uint32_t hour_offset, bool inject, cudaStream_t s);
void launch_recount(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_set_clock(Clock* clock, const Clock& value, cudaStream_t s);
void launch_commit(const Params& P, const DevPtrs& D, uint32_t hour_offset, bool lazy, bool zero_props, cudaStream_t s);
// tiles.cu: the plain movement hours with the grid tiles and their claims in shared memory
size_t tile_shared_bytes(const TileGeom& G);
size_t tile_sort_temp_bytes(uint32_t n);
cudaError_t build_tile_order(const Params& P, const DevPtrs& D, const TileGeom& G, const TilePtrs& TP, uint32_t* keys_a, uint32_t* keys_b, uint32_t* ids, void* temp,
                             size_t temp_bytes, cudaStream_t s);
void launch_count_housing(const Params& P, const DevPtrs& D, uint32_t* out, cudaStream_t s);
unsigned launch_hour_tiles(const Params& P, const DevPtrs& D, const TileGeom& G, const TilePtrs& TP, uint32_t n_generic_bound, uint32_t hour_offset, cudaStream_t s);
cudaError_t tiles_configure(const TileGeom& office, const TileGeom& house);
void launch_sleep(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s);
void launch_lock(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_unlock(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_vaccinate(const Params& P, const DevPtrs& D, uint64_t thr, uint32_t hour, cudaStream_t s);
void launch_import_state(const Params& P, const DevPtrs& D, const int32_t* cx, const int32_t* cy, uint32_t n_houses, uint32_t n_offices, uint32_t* bad, cudaStream_t s);
void launch_build_grid(const Params& P, const DevPtrs& D, uint32_t* collisions, cudaStream_t s);
// travel.cu: one cooperative launch each
cudaError_t launch_travel_leave(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, uint32_t* block_counts, TravelRecord* send,
                                uint32_t stride, unsigned grid_blocks, TravelRecord* const* peer_recv, uint32_t* const* peer_flags, uint32_t exchange_no, cudaStream_t s);
void launch_travel_recount(const Params& P, const DevPtrs& D, const TravelPtrs& T, uint32_t n_houses, uint32_t n_offices, cudaStream_t s);
cudaError_t launch_travel_arrive(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, const TravelRecord* recv, uint32_t stride,
                                 uint32_t n_houses, uint32_t n_offices, unsigned grid_blocks, const uint32_t* wait_flags, uint32_t exchange_no, cudaStream_t s);
cudaError_t launch_travel_exchange(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelPtrs& T, uint32_t* block_counts, TravelRecord* send, uint32_t stride,
                                   const TravelRecord* recv, uint32_t n_houses, uint32_t n_offices, unsigned grid_blocks, TravelRecord* const* peer_recv,
                                   uint32_t* const* peer_flags, const uint32_t* wait_flags, uint32_t exchange_no, cudaStream_t s);
unsigned travel_grid_blocks(int device, int per_sm);
}  // namespace epi
