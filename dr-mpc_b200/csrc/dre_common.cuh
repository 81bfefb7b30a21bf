// dre_common.cuh -- shared definitions for the DR-MPC B200 step engine (device + host).
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/dre.h"

#if defined(__CUDACC__)
#define DRE_HD __host__ __device__ __forceinline__
#define DRE_D __device__ __forceinline__
#else
#define DRE_HD inline
#define DRE_D inline
#endif

#if defined(__CUDACC__)
#define DRE_ALIGN16 __align__(16)
#else
#define DRE_ALIGN16 alignas(16)
#endif

#define DRE_RVO_EPSILON 0.00001f
#define DRE_PI 3.14159265358979323846

// done / info reason bits written to dre_step_out.done_bits (reference info['done'] sets,
// HA_and_PT env.py:327-331,391-404)
enum : uint32_t {
    DRE_DONE_COLLISION = 1u << 0,
    DRE_DONE_SAFETY_HUMAN = 1u << 1,
    DRE_DONE_ACTUATION = 1u << 2,
    DRE_DONE_TIMEOUT = 1u << 3,
    DRE_DONE_GOAL = 1u << 4,
    DRE_DONE_END_OF_PATH = 1u << 5,
    DRE_DONE_CORRIDOR_HIT = 1u << 6,
    DRE_DONE_SAFETY_CORRIDOR = 1u << 7,
    DRE_DONE_HA = 1u << 16,
    DRE_DONE_PT = 1u << 17,
    DRE_DONE_ANY = 1u << 18,
    DRE_DONE_FOR_REPLAY = 1u << 19,
};

// ---- Philox4x32-10 counter-based RNG (device-side goal resampling; SURVEY 8(d) "Seeds / RNG") ----
struct Philox {
    uint32_t c[4];
    uint32_t k[2];
};
DRE_HD void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
DRE_HD void philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4])
{
    uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// uniform double in [0,1) from two 32-bit words (53 bits)
DRE_HD double u01_from_bits(uint32_t hi, uint32_t lo)
{
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}
