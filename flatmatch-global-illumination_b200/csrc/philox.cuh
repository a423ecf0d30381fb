// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as
// easy as 1, 2, 3", SC'11; multipliers/Weyl constants as in Random123).  Replaces the reference's
// sequential libc rand() stream (photonmap.c:175-176,228; vector3_cl.c:107-108,131-132) and the
// OpenCL kernel's per-work-item LCG (photonmap.cl:21-25,272-275): every photon owns the
// sub-stream key = {seed, emitter}, counter = {photon lo, photon hi, event, 0}, so the sample
// set does not depend on how photons are scheduled over warps, SMs or GPUs.
#pragma once
#include <stdint.h>

namespace fmgi {

struct Philox4 { uint32_t w0, w1, w2, w3; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c1 = lo1; c3 = lo0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o = {c0, c1, c2, c3};
    return o;
}

// word -> xi in [0,1) with 24 random bits (exactly representable)
__host__ __device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
// the two low bytes the u24() conversions discard -> 16-bit roulette draw in [0,1)
__host__ __device__ __forceinline__ float r16(uint32_t a, uint32_t b)
{
    return (float)(((a & 255u) << 8) | (b & 255u)) * (1.0f / 65536.0f);
}

}  // namespace fmgi
