// Device helpers shared by the hour kernels (kernels.cu) and the traveller kernels (travel.cu).
#pragma once
#include <stdint.h>

#include "layout.h"

namespace epi {

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

__device__ __forceinline__ uint32_t count_category(uint32_t s) {
    const uint32_t st = s & ST_STATE_MASK;  // order of the CSV columns: S,E,I,H,R,D; 6 = empty slot (not counted)
    return st == ST_S ? 0u : st == ST_E ? 1u : st == ST_I ? ((s & ST_HOSP) ? 3u : 2u) : st == ST_R ? 4u : st == ST_D ? 5u : 6u;
}


// debugging timeline (EPI_TRACE=1): one stamp = (tag << 56) | low 56 bits of the GPU's nanosecond timer
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, unsigned tag, unsigned hour) {
    if (!trace) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned long long at = atomicAdd(trace, 2ull);
    if (at + 2ull < (1ull << 16)) { trace[1 + at] = ((unsigned long long)tag << 56) | (t & ((1ull << 56) - 1ull)); trace[2 + at] = hour; }
}

}  // namespace epi
