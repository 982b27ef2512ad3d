// Launchers of the sm_100a kernels in kernels.cu (callable from plain C++ translation units).
#pragma once
#include <cuda_runtime.h>

#include "layout.h"

namespace epi {
void launch_hospital_scan(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_hour(const Params& P, const DevPtrs& D, uint32_t hour_of_day, uint32_t hour_offset, bool inject, cudaStream_t s);
void launch_recount(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_set_clock(Clock* clock, const Clock& value, cudaStream_t s);
void launch_commit(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s);
void launch_sleep(const Params& P, const DevPtrs& D, uint32_t hour_offset, cudaStream_t s);
void launch_lock(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_unlock(const Params& P, const DevPtrs& D, cudaStream_t s);
void launch_vaccinate(const Params& P, const DevPtrs& D, uint64_t thr, uint32_t hour, cudaStream_t s);
void launch_build_grid(const Params& P, const DevPtrs& D, uint32_t* collisions, cudaStream_t s);
// travel.cu
void launch_travel_select(const Params& P, const DevPtrs& D, const TravelArgs& A, uint32_t* block_counts, uint32_t* total, uint32_t* out_slots,
                          uint32_t* out_dest, int phase, cudaStream_t s);
void launch_travel_pack(const Params& P, const DevPtrs& D, const uint32_t* send_slots, uint32_t n_send, TravelRecord* out, cudaStream_t s);
void launch_travel_install(const Params& P, const DevPtrs& D, const TravelArgs& A, const TravelRecord* in, uint32_t n_in, const uint32_t* in_slot,
                           const uint32_t* in_home, const uint32_t* in_work, cudaStream_t s);
void launch_travel_round(const Params& P, const DevPtrs& D, const TravelArgs& A, uint32_t n_in, uint32_t attempt, uint8_t* placed, const uint32_t* in_slot,
                         uint32_t* table_keys, uint32_t* table_vals, uint32_t table_mask, uint32_t* pending, cudaStream_t s);
}  // namespace epi
