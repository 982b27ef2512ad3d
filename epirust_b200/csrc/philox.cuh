// Philox4x32-10 counter-based RNG (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// Every draw on the hot path is philox(key = seed, counter = (agent, hour, block, domain)) -- no sequential RNG state,
// so any thread can compute any agent's draw (BASELINE.json north_star: "(seed, agent, hour, draw)").
// Replaces common::utils::RandomWrapper / rand::thread_rng (common/src/utils/random_wrapper.rs:23-35).
//
// Hour-step draws (DOM_STEP) -- one block serves the common agent-hour:
//   block 0: x = PICK (u32)  y = FACTOR (u32)  z,w = A (u64)
//   block 1: x = PX (u32)    y = PY (u32)
//   block 2+(j>>1): EXPOSE j (u64) = x,y for even j, z,w for odd j   (j = 0..7, Moore neighbour order)
// Injected-draw tables (epi_step_with_draws) carry 16 u64 per agent indexed by the slot numbers below; u32 draws take
// the low 32 bits.  All other domains use generic u64 slots: slot s = block s>>1, x,y for even s, z,w for odd s.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EPI_HD __host__ __device__ __forceinline__
#else
#define EPI_HD inline
#endif

namespace epi {

struct U4 {
    uint32_t x, y, z, w;
};

EPI_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

EPI_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

EPI_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

// draw domains (4th counter word)
enum : uint32_t { DOM_STEP = 0, DOM_INIT = 1, DOM_VACCINATE = 2, DOM_MIGRATE = 3, DOM_STARTINF = 4, DOM_ARRIVAL = 5 };
// hour-step slot numbers (index into an injected-draw row)
enum : uint32_t { SLOT_PICK = 0, SLOT_FACTOR = 1, SLOT_A = 2, SLOT_PX = 3, SLOT_PY = 4, SLOT_EXPOSE0 = 8 };

EPI_HD uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// generic u64 slot of a non-step domain
EPI_HD uint64_t philox_draw(uint64_t seed, uint32_t agent, uint32_t hour, uint32_t domain, uint32_t slot) {
    const U4 o = philox4x32_10(agent, hour, slot >> 1, domain, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (slot & 1u) ? u64_of(o.z, o.w) : u64_of(o.x, o.y);
}

// rand 0.8 `Rng::gen_bool(p)`: Bernoulli::new(p) -> p == 1.0 always true, else p_int = (p * 2^64) as u64, sample u64 < p_int.
EPI_HD uint64_t bernoulli_threshold(double p) {
    if (p >= 1.0) return UINT64_MAX;  // "always" sentinel
    if (p <= 0.0) return 0;
    return (uint64_t)(p * 18446744073709551616.0);
}
EPI_HD bool bernoulli(uint64_t draw, uint64_t thr) { return draw < thr || thr == UINT64_MAX; }

}  // namespace epi
