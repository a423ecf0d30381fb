// Philox counter-based generators (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
// 1, 2, 3", SC'11; multipliers/Weyl constants as in Random123).  They replace the reference's
// sequential libc rand() stream (photonmap.c:175-176,228; vector3_cl.c:107-108,131-132) and the
// OpenCL kernel's per-work-item LCG (photonmap.cl:21-25,272-275).
//
// The photon tracer draws Philox2x32-10 blocks: every event of a photon needs 64 random bits (two 24-bit
// uniforms and a 16-bit roulette draw), and with the seed as the only key word the ten round keys are
// kernel-wide constants, so a block is ten 32x32->64 multiplies and ten three-input XORs:
//     key     = seed
//     counter = {photon index bits 0..31, photon index bits 32..39 | emitter << 8 | event << 28}
//     event 15      emission position:  dx = u24(w0), dy = u24(w1)            (photonmap.c:175-176)
//     event 0       emission direction: xi1 = u24(w0), xi2 = u24(w1); roulette of bounce 1 = r16(w0, w1)
//     event b >= 1  direction after bounce b and the roulette of bounce b + 1, likewise
// The sample set therefore depends only on (seed, emitter, photon index) - not on how photons are scheduled
// over warps, SMs or GPUs.  Limits of the packed counter word: 2^40 photons per emitter, 2^20 emitters, 15
// bounces (fmgi_scene_trace refuses anything larger).  Philox4x32-10 is kept for its known-answer probe.
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

struct Philox2 { uint32_t w0, w1; };

constexpr uint32_t kPhiloxM2 = 0xD256D193u, kPhiloxW = 0x9E3779B9u;
constexpr uint32_t kEventEmitPosition = 15u;
constexpr int kPhiloxMaxDepth = 15, kPhiloxMaxEmitters = 1 << 20;
constexpr unsigned long long kPhiloxMaxPhotons = 1ull << 40;

// Philox2x32-10 with the round keys key + r * W supplied by the caller (kernel-wide constants).
__host__ __device__ __forceinline__ Philox2 philox2x32_10(uint32_t c0, uint32_t c1, const uint32_t *round_keys)
{
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        const uint32_t hi = __umulhi(kPhiloxM2, c0), lo = kPhiloxM2 * c0;
#else
        const uint64_t prod = (uint64_t)kPhiloxM2 * c0;
        const uint32_t hi = (uint32_t)(prod >> 32), lo = (uint32_t)prod;
#endif
        c0 = hi ^ round_keys[r] ^ c1;
        c1 = lo;
    }
    Philox2 o = {c0, c1};
    return o;
}

// the packed counter word of a photon event
__host__ __device__ __forceinline__ uint32_t philox_event_word(uint32_t photon_hi, uint32_t emitter, uint32_t event)
{
    return photon_hi | (emitter << 8) | (event << 28);
}

// word -> xi in [0,1) with 24 random bits (exactly representable)
__host__ __device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
// the two low bytes the u24() conversions discard -> 16-bit roulette draw in [0,1)
__host__ __device__ __forceinline__ float r16(uint32_t a, uint32_t b)
{
    return (float)(((a & 255u) << 8) | (b & 255u)) * (1.0f / 65536.0f);
}

}  // namespace fmgi
