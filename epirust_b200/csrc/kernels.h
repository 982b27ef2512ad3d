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
}  // namespace epi
