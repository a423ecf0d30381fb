// Device code of the photon-mapping hot path: emit -> closest hit -> deposit -> re-emit.
// Written for sm_100a (B200).  Restates, from scratch, what the reference does in
// photonmap.cl:161-281 (OpenCL kernel) / photonmap.c:164-272 (native), with the native path's
// semantics where the two differ.
//
// Shape of the computation (see DESIGN.md for the reasoning and the measurements):
//   * one persistent kernel for the whole bake: every lane carries one photon; when a photon
//     dies (miss, or last allowed bounce) the warp's dead lanes are found with a ballot and
//     refilled from the warp's chunk of the global photon index space, so the closest-hit loop
//     always runs with full warps ("wavefront" compaction done in registers);
//   * the rectangle soup lives in shared memory as plane-grouped axis-aligned records; every
//     lane of a warp reads the same record (broadcast), the loop trip count is warp uniform;
//   * back-face culling (rectangle.c:70-72) costs nothing per test: for a group whose normal the
//     ray cannot face, the lane's reciprocal direction is replaced by NaN, which fails the
//     single unsigned compare that also implements 0 <= t < best;
//   * deposits are one 16-byte vector reduction (RED.E.ADD.F32x4) per bounce into the L2-resident
//     atlas, optionally warp-aggregated with __match_any_sync;
//   * per-photon Philox4x32-10 sub-streams (philox.cuh) replace the sequential libc stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"
#include "scene_tables.h"

namespace fmgi {

constexpr int kTraceThreads = 256;
constexpr int kChunkPhotons = 256;          // photon indices a warp claims per global atomic
constexpr unsigned kFullMask = 0xffffffffu;

struct TraceParams {
    // closest-hit tables (global memory; staged into shared memory by the soup kernel)
    const float4 *axis;          // 2 float4 per AxisRect, grouped
    const float4 *general;       // 4 float4 per GeneralRect
    int group_begin[kNumAxisGroups + 1];
    int num_general;
    // shading tables
    const float4 *shade;         // 6 float4 per wall
    const float4 *emitters;      // 6 float4 per emitter
    // this shard's photon index space: jobs [job_begin[e], job_begin[e+1]) belong to emitter e and
    // map to photon indices photon_first[e] + (job - job_begin[e])
    const unsigned long long *job_begin;
    const unsigned long long *photon_first;
    int num_emitters;
    unsigned long long total_jobs;
    unsigned long long *work_counter;
    // output
    float4 *atlas;
    unsigned long long *counters;   // photons, rays, deposits, mirror bounces
    int32_t *path_out;              // probe builds only
    int max_depth;
    uint32_t seed;
};

// ---- closest hit against the shared-memory soup ----------------------------------------------

struct HitRec { float t; int slot; };   // slot: >= 0 position in the axis table, < -1: ~position in general table, -1 miss

// One (axis, sign) group.  a = 1/d[k] if the ray can face the group else NaN; b = -o[k]*a.
__device__ __forceinline__ void scan_axis_group(const float4 *__restrict__ tab, int begin, int end,
                                                float a, float b, float oi, float di, float oj, float dj,
                                                float &best, int &slot)
{
#pragma unroll 4
    for (int r = begin; r < end; r++) {
        const float4 q0 = tab[2 * r];
        const float4 q1 = tab[2 * r + 1];
        const float t = fmaf(q0.x, a, b);
        const float pi = fmaf(t, di, oi);
        const float pj = fmaf(t, dj, oj);
        // 0 <= t < best in one compare: negative and NaN floats are large unsigned integers
        const bool ok = (__float_as_uint(t) < __float_as_uint(best)) &&
                        pi >= q0.y && pi <= q0.z && pj >= q0.w && pj <= q1.x;
        if (ok) { best = t; slot = r; }
    }
}

// rectangle.c:67-95 for an arbitrarily oriented rectangle
__device__ __forceinline__ void scan_general(const float4 *__restrict__ tab, int count,
                                             float ox, float oy, float oz, float dx, float dy, float dz,
                                             float &best, int &slot)
{
    for (int r = 0; r < count; r++) {
        const float4 g0 = tab[4 * r], g1 = tab[4 * r + 1], g2 = tab[4 * r + 2], g3 = tab[4 * r + 3];
        const float denom = g0.x * dx + g0.y * dy + g0.z * dz;
        const float num = g0.w - (g0.x * ox + g0.y * oy + g0.z * oz);
        const float t = __fdividef(num, denom);
        const float ex = fmaf(t, dx, ox) - g3.x, ey = fmaf(t, dy, oy) - g3.y, ez = fmaf(t, dz, oz) - g3.z;
        const float u = g1.x * ex + g1.y * ey + g1.z * ez;
        const float v = g2.x * ex + g2.y * ey + g2.z * ez;
        const bool ok = denom < 0.0f && (__float_as_uint(t) < __float_as_uint(best)) &&
                        u >= 0.0f && v >= 0.0f && u <= g1.w && v <= g2.w;
        if (ok) { best = t; slot = ~r - 1; }     // -2, -3, ...
    }
}

struct SoupTables {
    const float4 *axis;
    const float4 *general;
    int group_begin[kNumAxisGroups + 1];
    int num_general;
};

// Closest front-facing hit.  Returns the wall index (or -1) and the distance recomputed from the
// winning plane as (c - o[k]) / d[k], the reference's formulation for an axis-parallel normal.
__device__ __forceinline__ int closest_hit_soup(const SoupTables &s, float ox, float oy, float oz,
                                                float dx, float dy, float dz, float &t_out)
{
    const float nanv = __int_as_float(0x7fc00000);
    const float ix = __frcp_rn(dx), iy = __frcp_rn(dy), iz = __frcp_rn(dz);
    float best = __int_as_float(0x7f800000);
    int slot = -1;
    // group 2k: normal +k, hit by rays with d[k] < 0; group 2k+1: normal -k, d[k] > 0
    {
        const float a0 = dx < 0.0f ? ix : nanv, a1 = dx > 0.0f ? ix : nanv;
        scan_axis_group(s.axis, s.group_begin[0], s.group_begin[1], a0, -ox * a0, oy, dy, oz, dz, best, slot);
        scan_axis_group(s.axis, s.group_begin[1], s.group_begin[2], a1, -ox * a1, oy, dy, oz, dz, best, slot);
    }
    {
        const float a0 = dy < 0.0f ? iy : nanv, a1 = dy > 0.0f ? iy : nanv;
        scan_axis_group(s.axis, s.group_begin[2], s.group_begin[3], a0, -oy * a0, ox, dx, oz, dz, best, slot);
        scan_axis_group(s.axis, s.group_begin[3], s.group_begin[4], a1, -oy * a1, ox, dx, oz, dz, best, slot);
    }
    {
        const float a0 = dz < 0.0f ? iz : nanv, a1 = dz > 0.0f ? iz : nanv;
        scan_axis_group(s.axis, s.group_begin[4], s.group_begin[5], a0, -oz * a0, ox, dx, oy, dy, best, slot);
        scan_axis_group(s.axis, s.group_begin[5], s.group_begin[6], a1, -oz * a1, ox, dx, oy, dy, best, slot);
    }
    if (s.num_general)
        scan_general(s.general, s.num_general, ox, oy, oz, dx, dy, dz, best, slot);

    int id = -1;
    t_out = best;
    if (slot >= 0) {
        const float4 q0 = s.axis[2 * slot];
        const float4 q1 = s.axis[2 * slot + 1];
        id = __float_as_int(q1.y);
        const int k = slot < s.group_begin[2] ? 0 : (slot < s.group_begin[4] ? 1 : 2);
        const float ok = k == 0 ? ox : (k == 1 ? oy : oz);
        const float dk = k == 0 ? dx : (k == 1 ? dy : dz);
        t_out = __fdiv_rn(__fsub_rn(q0.x, ok), dk);
    } else if (slot < -1) {
        // rectangle.c:70-75 for the winner: t = n.(pos - o) / n.d with IEEE operations
        const float4 *g = s.general + 4 * (~(slot + 1));
        const float4 g0 = g[0], g3 = g[3];
        id = __float_as_int(g3.w);
        const float denom = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, dx), __fmul_rn(g0.y, dy)), __fmul_rn(g0.z, dz));
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, __fsub_rn(g3.x, ox)), __fmul_rn(g0.y, __fsub_rn(g3.y, oy))),
                                    __fmul_rn(g0.z, __fsub_rn(g3.z, oz)));
        t_out = __fdiv_rn(num, denom);
    }
    return id;
}

// ---- texel index: rectangle.c:205-230, same operations in the same order, no contraction ----------

__device__ __forceinline__ int tile_index(const float4 q0, const float4 q1, const float4 q2, int tiles,
                                          float px, float py, float pz)
{
    const float ex = __fsub_rn(px, q0.x), ey = __fsub_rn(py, q0.y), ez = __fsub_rn(pz, q0.z);
    const float du = __fadd_rn(__fadd_rn(__fmul_rn(q1.x, ex), __fmul_rn(q1.y, ey)), __fmul_rn(q1.z, ez));
    const float dv = __fadd_rn(__fadd_rn(__fmul_rn(q2.x, ex), __fmul_rn(q2.y, ey)), __fmul_rn(q2.z, ez));
    const int tw = tiles & 0xffff, th = tiles >> 16;
    int tx = __float2int_rz(__fdiv_rn(__fmul_rn(du, (float)tw), q1.w));
    int ty = __float2int_rz(__fdiv_rn(__fmul_rn(dv, (float)th), q2.w));
    tx = min(max(tx, 0), tw - 1);
    ty = min(max(ty, 0), th - 1);
    return __float_as_int(q0.w) + ty * tw + tx;
}

// ---- hemisphere sampling: vector3_cl.c:102-149 -------------------------------------------------------

// Malley disk sample around n with the precomputed basis (u, v).  sqrt and the phi product round
// like the reference's double-then-float evaluation; sin/cos use the SFU.
__device__ __forceinline__ void sample_hemisphere(float xi1, float xi2, bool sky, const float4 n, const float4 u,
                                                  const float4 v, float &dx, float &dy, float &dz)
{
    const float r = __fsqrt_rn(xi1);
    const float phi = __fmul_rn(2.0f * 3.141592f, xi2);
    float sp, cp;
    __sincosf(phi, &sp, &cp);
    float a = r * cp;
    const float b = r * sp;
    const float c = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(r, r)));
    if (sky) a = fabsf(a);                 // vector3_cl.c:115-116
    dx = fmaf(n.x, c, fmaf(v.x, b, u.x * a));
    dy = fmaf(n.y, c, fmaf(v.y, b, u.y * a));
    dz = fmaf(n.z, c, fmaf(v.z, b, u.z * a));
}

// ---- deposit ------------------------------------------------------------------------------------------

template <int kDeposit>
__device__ __forceinline__ void deposit(float4 *__restrict__ atlas, int idx, float r, float g, float b, bool active)
{
    if (kDeposit == FMGI_DEPOSIT_WARP_AGG) {
        // Lanes that target the same texel elect a leader that adds the group's sum once.
        const unsigned act = __ballot_sync(kFullMask, active);
        if (active) {
            const unsigned peers = __match_any_sync(act, idx);
            const int leader = __ffs(peers) - 1;
            const int lane = threadIdx.x & 31;
            if (peers != (1u << lane)) {
                // segmented sum over the (usually tiny) peer set
                unsigned rest = peers & ~(1u << leader);
                float sr = r, sg = g, sb = b;
                // every peer must take part in the shuffles with the same mask
                while (rest) {
                    const int src = __ffs(rest) - 1;
                    const float orr = __shfl_sync(peers, r, src);
                    const float og = __shfl_sync(peers, g, src);
                    const float ob = __shfl_sync(peers, b, src);
                    sr += orr; sg += og; sb += ob;
                    rest &= rest - 1;
                }
                r = sr; g = sg; b = sb;
            }
            if (lane == leader)
                atomicAdd(atlas + idx, make_float4(r, g, b, 0.0f));
        }
    } else if (kDeposit == FMGI_DEPOSIT_SCALAR) {
        if (active) {
            float *t = reinterpret_cast<float *>(atlas + idx);
            atomicAdd(t + 0, r); atomicAdd(t + 1, g); atomicAdd(t + 2, b);
        }
    } else {
        if (active)
            atomicAdd(atlas + idx, make_float4(r, g, b, 0.0f));
    }
}

}  // namespace fmgi
