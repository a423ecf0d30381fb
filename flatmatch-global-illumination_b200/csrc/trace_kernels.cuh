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
//   * scenes above 64 colliders whose colliders are all axis parallel (example.png included: everything
//     parseLayout.c emits) find the closest hit by walking a box decomposition of the flat in which every
//     collider lies on a box face (rooms_walk, room_tables.h): a ray leaves its box through the nearest of the
//     three faces ahead - one 256-bit load, 25 instructions - and the face's code says what is there: a wall (the
//     hit), the next box, nothing, or a grid record for a face that holds several things; the photon loop
//     interleaves these box steps with the shading of the lanes that already hit something;
//   * scenes with arbitrarily oriented colliders use a floor-plan grid (GridWalk): a 2-D DDA through per-cell,
//     per-sign-combination wall lists whose loop is 35 instructions of straight-line predicated PTX - test the
//     pending record, step, fetch with ONE 256-bit load that also carries the continuation - then one head lookup
//     per z plane crossed before the wall hit;
//   * a bare room keeps the brute-force soup in shared memory: per-axis lists of axis-parallel
//     records, warp-uniform trip count, broadcast loads; back-face culling (rectangle.c:70-72)
//     halves the work instead of costing a test: the records of one normal axis are split by normal
//     sign into two interleaved lists and every lane walks only the list its ray can face; its
//     horizontal rectangles go through the grid's plane tables when they fit;
//   * 0 <= t < best is one unsigned compare (negative and NaN floats are large unsigned
//     integers); containment is |p - mid| <= half, which moves half of the compare work from the
//     ALU pipe to the FMA pipe (the ALU pipe was the top pipe of the first version, see profiles/);
//   * deposits are one 16-byte vector reduction (RED.E.ADD.F32x4) per bounce into the atlas,
//     optionally warp-aggregated with __match_any_sync (measured: no gain, not the default);
//   * per-photon Philox2x32-10 sub-streams (philox.cuh) replace the sequential libc stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"
#include "scene_tables.h"

namespace fmgi {

constexpr int kTraceThreads = 256;
constexpr int kChunkPhotons = 256;          // photon indices a warp claims per global atomic (big bakes)
constexpr int kMinChunkPhotons = 32;        // ... small bakes use smaller chunks so that every SM gets work
constexpr unsigned kFullMask = 0xffffffffu;
// internal kernel variant: soup tier whose horizontal rectangles go through the grid's plane tables
constexpr int kTierSoupPlanes = 3;
// room tier (room_tables.h): FMGI_TIER_ROOMS
constexpr int kTierRooms = FMGI_TIER_ROOMS;
constexpr int kRoomBoxVec = 16;             // float4 per RoomBox (room_tables.h)

struct TraceParams {
    // closest-hit tables (global memory; staged into shared memory by the soup kernel)
    const float4 *axis;          // 3 float4 per AxisPairBlock, 2 blocks (P, M) per pair
    const float4 *general;       // 4 float4 per GeneralRect
    int pair_begin[4];           // pairs of axis k: pair_begin[k] .. pair_begin[k+1]
    int num_general;
    // grid tier (scene_tables.h): list heads, then the lists' remaining records
    const float4 *grid_table;    // 2 float4 per GridRec
    GridDesc grid;
    // shading tables
    const float4 *shade;         // 6 float4 per wall
    const float4 *emitters;      // 6 float4 per emitter
    const float *ao_width;       // ambient occlusion only: the walls' width / height vectors (3 floats each)
    const float *ao_height;
    // this shard's photon index space in chunks of `chunk` photons of ONE emitter (the unit a warp
    // claims): chunks [job_begin[e], job_begin[e+1]) belong to emitter e; chunk j of emitter e covers
    // photon indices photon_first[e] + [j * chunk, min((j + 1) * chunk, photon_count[e]))
    const unsigned long long *job_begin;
    const unsigned long long *photon_first;
    const unsigned long long *photon_count;
    int num_emitters;
    int chunk;                       // photons per chunk: kMinChunkPhotons .. kChunkPhotons
    unsigned long long total_jobs;   // chunks
    unsigned long long *work_counter;
    // output
    float4 *atlas;
    unsigned long long *counters;   // photons, rays, deposits, mirror bounces, (work counter), rectangle tests (grid tier)
    int32_t *path_out;              // probe builds only
    int max_depth;
    uint32_t seed;
    uint32_t philox_keys[10];       // seed + r * W: the Philox2x32 round keys of this bake (philox.cuh)
    int grid_has_misc;              // the walk lists hold misc records (scene_tables.h)
    int one;                        // 1, opaque to the compiler: x * one + y keeps integer updates on the FMA pipe
    // room tier (room_tables.h): boxes (16 float4 each: 8 octant records), face grids and box bounds (2 float4 each),
    // kd-tree nodes for point location (1 float4 each)
    const float4 *room_boxes;
    const float4 *room_face_grids;  // RoomFaceGrid records (2 float4 each)
    const unsigned *room_face_cells; // the codes they index
    const float4 *room_bounds;
    const float4 *room_nodes;
    const int2 *room_starts;        // per emitter: {code, normal axis} (RoomStart)
    float room_lo[3], room_hi[3];   // the root box
    // table sizes: read only by the bounds-checked build (FMGI_CHECKED, lib/libfmgi_cuda_checked.so)
    unsigned grid_records, num_walls, num_texels, room_num_boxes, room_num_face_grids, room_num_face_cells, room_num_nodes;
};

// ---- bounds-checked build ----------------------------------------------------------------------------------------
// compute-sanitizer is not available on the GPU pool, so the library can be built with -DFMGI_CHECKED: every
// data-dependent index - grid table records, shading / emitter records, atlas texels - is then compared with its
// table size before it is used; a violation is counted (counters[6], first site code in counters[7]), the access
// is skipped, and fmgi_stats.bounds_violations reports the count (the Python mirror raises on any).  The GPU test
// suite and profiles/sanitize_small.py run against that library as the memcheck stand-in (tools/gpu_checked.sh).
#ifdef FMGI_CHECKED
__device__ __noinline__ void fmgi_report(const TraceParams &p, int code)
{
    atomicAdd(p.counters + 6, 1ull);
    atomicCAS(p.counters + 7, 0ull, (unsigned long long)code);
}
__device__ __forceinline__ bool fmgi_check(const TraceParams &p, bool ok, int code)
{
    if (!ok) fmgi_report(p, code);
    return ok;
}
#define FMGI_CHECK(p, cond, code) fmgi_check((p), (cond), (code))
#else
#define FMGI_CHECK(p, cond, code) (true)
#endif

// ---- closest hit against the shared-memory soup ----------------------------------------------

// Hit code: >= 0: 2 * (global pair index) + (0: rectangle a, 1: rectangle b) of the list the ray
// faces on that axis; <= -2: -(index into the general table) - 2; -1: miss.

// One axis-parallel rectangle test, written in PTX so that it stays the ten instructions it needs:
// three FFMA + two FADD on the FMA pipe, one unsigned and two float compares chained through one
// predicate, and two selects on the ALU pipe.
//   t  = c * a + b                      (ray parameter on the plane)
//   pi = t * di + oi - mid_i, pj alike  (hit point relative to the rectangle centre)
//   ok = (0 <= t < best) && |pi| <= half_i && |pj| <= half_j
__device__ __forceinline__ void axis_test(float c, float mid_i, float half_i, float mid_j, float half_j,
                                          float a, float b, float oi, float di, float oj, float dj,
                                          int cur, float &best, int &code)
{
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .f32 t, pi, pj;\n\t"
        ".reg .b32 tb, bb;\n\t"
        "fma.rn.f32 t, %2, %7, %8;\n\t"
        "fma.rn.f32 pi, t, %10, %9;\n\t"
        "fma.rn.f32 pj, t, %12, %11;\n\t"
        "sub.rn.f32 pi, pi, %3;\n\t"
        "sub.rn.f32 pj, pj, %5;\n\t"
        "abs.f32 pi, pi;\n\t"
        "abs.f32 pj, pj;\n\t"
        "mov.b32 tb, t;\n\t"
        "mov.b32 bb, %0;\n\t"
        "setp.lt.u32 p, tb, bb;\n\t"
        "setp.le.and.f32 p, pi, %4, p;\n\t"
        "setp.le.and.f32 p, pj, %6, p;\n\t"
        "selp.f32 %0, t, %0, p;\n\t"
        "selp.b32 %1, %13, %1, p;\n\t"
        "}"
        : "+f"(best), "+r"(code)
        : "f"(c), "f"(mid_i), "f"(half_i), "f"(mid_j), "f"(half_j), "f"(a), "f"(b), "f"(oi), "f"(di), "f"(oj),
          "f"(dj), "r"(cur));
}

// One normal axis.  blk points at the lane's own first block (P or M list); a = 1/d[k], b = -o[k]*a.
__device__ __forceinline__ void scan_axis(const float4 *__restrict__ blk, int pair0, int pair1,
                                          float a, float b, float oi, float di, float oj, float dj,
                                          float &best, int &code)
{
    blk += 6 * pair0;
#pragma unroll 4
    for (int j = pair0; j < pair1; j++, blk += 6) {
        const float4 qa = blk[0];
        const float4 qb = blk[1];
        const float4 qc = blk[2];
        axis_test(qa.x, qa.y, qa.z, qa.w, qc.x, a, b, oi, di, oj, dj, 2 * j, best, code);
        axis_test(qb.x, qb.y, qb.z, qb.w, qc.y, a, b, oi, di, oj, dj, 2 * j + 1, best, code);
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
        if (ok) { best = t; slot = -r - 2; }
    }
}

struct SoupTables {
    const float4 *axis;
    const float4 *general;
    int pair_begin[4];
    int num_general;
};

// Wall index of the soup winner `code` (-1: miss) and its distance recomputed from the winning plane
// as (c - o[k]) / d[k], the reference's formulation for an axis-parallel normal.
__device__ __forceinline__ int soup_winner(const SoupTables &s, int code, float best, float ox, float oy, float oz,
                                           float dx, float dy, float dz, float &t_out)
{
    int id = -1;
    t_out = best;
    if (code >= 0) {
        const int pair = code >> 1;
        const int k = pair < s.pair_begin[1] ? 0 : (pair < s.pair_begin[2] ? 1 : 2);
        const float ok = k == 0 ? ox : (k == 1 ? oy : oz);
        const float dk = k == 0 ? dx : (k == 1 ? dy : dz);
        const float4 *blk = s.axis + 6 * pair + (dk > 0.0f ? 3 : 0);
        const float c = (code & 1) ? blk[1].x : blk[0].x;
        id = __float_as_int((code & 1) ? blk[2].w : blk[2].z);
        t_out = __fdiv_rn(__fsub_rn(c, ok), dk);
    } else if (code < -1) {
        // rectangle.c:70-75 for the winner: t = n.(pos - o) / n.d with IEEE operations
        const float4 *g = s.general + 4 * (-code - 2);
        const float4 g0 = g[0], g3 = g[3];
        id = __float_as_int(g3.w);
        const float denom = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, dx), __fmul_rn(g0.y, dy)), __fmul_rn(g0.z, dz));
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, __fsub_rn(g3.x, ox)), __fmul_rn(g0.y, __fsub_rn(g3.y, oy))),
                                    __fmul_rn(g0.z, __fsub_rn(g3.z, oz)));
        t_out = __fdiv_rn(num, denom);
    }
    return id;
}

// Closest front-facing hit.  Returns the wall index (or -1) and the distance recomputed from the
// winning plane as (c - o[k]) / d[k], the reference's formulation for an axis-parallel normal.
__device__ __forceinline__ int closest_hit_soup(const SoupTables &s, float ox, float oy, float oz,
                                                float dx, float dy, float dz, float &t_out)
{
    const float nanv = __int_as_float(0x7fc00000);
    float best = __int_as_float(0x7f800000);
    int code = -1;
    // a lane whose d[k] is exactly zero can face neither list of axis k
    const float ax = dx != 0.0f ? __frcp_rn(dx) : nanv;
    const float ay = dy != 0.0f ? __frcp_rn(dy) : nanv;
    const float az = dz != 0.0f ? __frcp_rn(dz) : nanv;
    // list P (normal +k) is block 0 of a pair, list M (normal -k) block 1; d[k] > 0 faces M
    const int sx = dx > 0.0f ? 3 : 0, sy = dy > 0.0f ? 3 : 0, sz = dz > 0.0f ? 3 : 0;
    scan_axis(s.axis + sx, s.pair_begin[0], s.pair_begin[1], ax, -ox * ax, oy, dy, oz, dz, best, code);
    scan_axis(s.axis + sy, s.pair_begin[1], s.pair_begin[2], ay, -oy * ay, ox, dx, oz, dz, best, code);
    scan_axis(s.axis + sz, s.pair_begin[2], s.pair_begin[3], az, -oz * az, ox, dx, oy, dy, best, code);
    if (s.num_general)
        scan_general(s.general, s.num_general, ox, oy, oz, dx, dy, dz, best, code);

    return soup_winner(s, code, best, ox, oy, oz, dx, dy, dz, t_out);
}

// ---- closest hit through the floor-plan grid (grid tier) ---------------------------------------------

__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 ldg2(const float4 *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
// One 256-bit read-only load (sm_100: LDG.E.256) of a 32-byte aligned pair of float4: one request to
// L1 instead of two.  The trace kernel's lanes read scattered 32-byte records, which made the L1
// data pipe (wavefronts, one per distinct line and request) the busiest unit of the grid tier.
__device__ __forceinline__ void ldg256(const float4 *p, float4 &a, float4 &b)
{
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// Rare walk-list records (tag bit 30): a horizontal rectangle beyond the plane table or an arbitrarily
// oriented rectangle.  These sit in all four walk lists, so facing is tested here.  Returns the ray
// parameter of a valid hit closer than `best`, else +inf (by value: the caller's running minimum
// stays in registers).
__device__ __noinline__ float grid_test_misc(const float4 *__restrict__ general, float4 q0, float c, unsigned tag, float ox,
                                             float oy, float oz, float dx, float dy, float dz, float best)
{
    const float inf = __int_as_float(0x7f800000);
    if (tag & kTagHorizontal) {
        // rectangle.c:67-95 for an arbitrarily oriented rectangle
        const float4 *g = general + 4 * (tag & kTagIdMask);
        const float4 g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + 2), g3 = __ldg(g + 3);
        const float denom = g0.x * dx + g0.y * dy + g0.z * dz;
        const float num = g0.w - (g0.x * ox + g0.y * oy + g0.z * oz);
        const float t = __fdividef(num, denom);
        const float ex = fmaf(t, dx, ox) - g3.x, ey = fmaf(t, dy, oy) - g3.y, ez = fmaf(t, dz, oz) - g3.z;
        const float u = g1.x * ex + g1.y * ey + g1.z * ez;
        const float v = g2.x * ex + g2.y * ey + g2.z * ez;
        const bool ok = denom < 0.0f && (__float_as_uint(t) < __float_as_uint(best)) && u >= 0.0f && v >= 0.0f &&
                        u <= g1.w && v <= g2.w;
        return ok ? t : inf;
    }
    const bool facing = (tag & kTagNegative) ? dz > 0.0f : dz < 0.0f;    // back-face culling, rectangle.c:70-72
    const float t = __fdividef(c - oz, dz);
    const float pi = fmaf(t, dx, ox) - q0.x;
    const float pj = fmaf(t, dy, oy) - q0.z;
    const bool ok = facing && (__float_as_uint(t) < __float_as_uint(best)) && fabsf(pi) <= q0.y && fabsf(pj) <= q0.w;
    return ok ? t : inf;
}

// Closest-hit query through the grid.  (Shapes that were measured and dropped: a kernel that interleaved
// the walk steps of different rays - lanes whose walk was over waited until most of the warp was
// waiting, then shaded together; and, in a CPU replay of the warp, walks capped at M steps per round
// with unfinished lanes carried over - the shading phase then runs with fewer lanes and eats the gain.)
struct GridWalk {
    float best;            // ray parameter of the best hit so far (+inf: none)
    int win;               // index in T of the best hit's record, -1: none

    __device__ __forceinline__ void reset() { best = __int_as_float(0x7f800000); win = -1; }

    // Horizontal planes the ray can face - one head lookup per plane at the crossing point, nearest
    // plane first (the table is sorted), so that a hit bounds the remaining planes away.  Only planes
    // crossed before the current best hit are looked up.  The first kFastPlanes planes of the ray's
    // direction (a flat has two or three: floor + sills, ceiling + door and window lintels) are handled
    // in straight-line code: all crossing points first, then all head loads back to back - independent,
    // so their latencies overlap instead of adding up (the sequential loop spent 27 % of the kernel's
    // stall samples waiting for one head after the other, profiles/) - then the containment tests in
    // order.  Further planes take the loop.
    static constexpr int kFastPlanes = 3;

    // the head's rectangle does not contain the crossing point: try the rest of the cell's list [q, end)
    __device__ __forceinline__ bool plane_rest(const TraceParams &p, int q, int end, float x, float y, float t,
                                               unsigned &tests, bool count)
    {
        for (; q < end; q++) {
            if (!FMGI_CHECK(p, (unsigned)q < p.grid_records, 2)) break;
            const float4 q0 = __ldg(p.grid_table + 2 * q);
            if (count) tests++;
            if (fabsf(x - q0.x) <= q0.y && fabsf(y - q0.z) <= q0.w) { best = t; win = q; return true; }
        }
        return false;
    }

    // The lookup is written as ONE predicated instruction stream for the first three planes of the ray's direction:
    // ray parameters, crossing points, cells and head indices of all three, their head loads back to back
    // (independent, so the latencies overlap), then the containment tests in plane order with best / win updated
    // by select.  The first version branched per plane (head test, "rest of the list" loop, break out), which cost
    // 7.5 warp instructions per ray at 19 of 32 lanes - a fifth of the kernel (profiles/r2a_example_sections.txt).
    // What is left to a branch is the rest of a cell's list when the head's rectangle does not contain the crossing
    // point (a cell that straddles two rooms: one ray in five on example.png): ONE loop serves all three planes.
    template <bool kCount>
    __device__ __forceinline__ void planes(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                           float dz, unsigned &tests)
    {
        const GridDesc &g = p.grid;
        const int dir = dz < 0.0f ? 0 : 1;                             // d.z < 0 faces normals +z
        const float iz = rcp_fast(dz);                                 // d.z == 0: t = +-inf or NaN, never < best
        const float4 fz = *reinterpret_cast<const float4 *>(g.fast_z[dir]);
        const int4 fb = *reinterpret_cast<const int4 *>(g.fast_base[dir]);
        const float zz[kFastPlanes] = {fz.x, fz.y, fz.z};
        const int bb[kFastPlanes] = {fb.x, fb.y, fb.z};
        float tt[kFastPlanes], xx[kFastPlanes], yy[kFastPlanes];
        int hh[kFastPlanes];
        bool go[kFastPlanes];
        float4 h0[kFastPlanes], h1[kFastPlanes];
#pragma unroll
        for (int i = 0; i < kFastPlanes; i++) {
            tt[i] = (zz[i] - oz) * iz;
            xx[i] = fmaf(tt[i], dx, ox); yy[i] = fmaf(tt[i], dy, oy);
            const int px = __float2int_rd(fmaf(xx[i], g.inv_cell, g.bx)), py = __float2int_rd(fmaf(yy[i], g.inv_cell, g.by));
            go[i] = (__float_as_uint(tt[i]) < __float_as_uint(best)) && (unsigned)px < (unsigned)g.nx &&
                    (unsigned)py < (unsigned)g.ny;
            hh[i] = bb[i] + py * g.nx + px;
            go[i] = go[i] && FMGI_CHECK(p, (unsigned)hh[i] < p.grid_records, 1);
        }
#pragma unroll
        for (int i = 0; i < kFastPlanes; i++) {
            h0[i] = make_float4(0.0f, -1.0f, 0.0f, -1.0f);
            h1[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (go[i]) ldg256(p.grid_table + 2 * hh[i], h0[i], h1[i]);
        }
        unsigned need = 0;
#pragma unroll
        for (int i = 0; i < kFastPlanes; i++) {
            // a hit on a nearer plane makes the farther ones moot (tt[i] >= best)
            const bool live = go[i] && (__float_as_uint(tt[i]) < __float_as_uint(best));
            const bool in = live && fabsf(xx[i] - h0[i].x) <= h0[i].y && fabsf(yy[i] - h0[i].z) <= h0[i].w;
            if (kCount) tests += live && h0[i].y >= 0.0f ? 1u : 0u;    // dummy heads (half width -1) are not tests
            best = in ? tt[i] : best;
            win = in ? hh[i] : win;
            need |= (live && !in && __float_as_int(h1[i].z) < __float_as_int(h1[i].w)) ? (1u << i) : 0u;
        }
        const unsigned act = __activemask();
        if (__any_sync(act, need != 0u)) {
#pragma unroll
            for (int i = 0; i < kFastPlanes; i++)
                if ((need >> i & 1u) && (__float_as_uint(tt[i]) < __float_as_uint(best)))
                    plane_rest(p, __float_as_int(h1[i].z), __float_as_int(h1[i].w), xx[i], yy[i], tt[i], tests, kCount);
        }
        const bool down = dir == 0;
        const int count = down ? g.planes_up : g.planes_down;
        const int base = down ? 0 : g.down_base;
#pragma unroll 1
        for (int pl = kFastPlanes; pl < g.planes_max; pl++) {
            const float z = down ? g.plane_z[pl] : g.plane_z[kMaxPlanesPerSign + pl];
            const float t = (z - oz) * iz;
            const float x = fmaf(t, dx, ox), y = fmaf(t, dy, oy);
            const int px = __float2int_rd(fmaf(x, g.inv_cell, g.bx)), py = __float2int_rd(fmaf(y, g.inv_cell, g.by));
            const bool go2 = pl < count && (__float_as_uint(t) < __float_as_uint(best)) && (unsigned)px < (unsigned)g.nx &&
                             (unsigned)py < (unsigned)g.ny;
            if (go2 && FMGI_CHECK(p, (unsigned)(base + pl * g.ncell + py * g.nx + px) < p.grid_records, 3)) {
                const int head = base + pl * g.ncell + py * g.nx + px;
                float4 q0, q1;
                ldg256(p.grid_table + 2 * head, q0, q1);
                if (kCount) tests += q0.y >= 0.0f ? 1u : 0u;
                if (fabsf(x - q0.x) <= q0.y && fabsf(y - q0.z) <= q0.w) { best = t; win = head; }
                else plane_rest(p, __float_as_int(q1.z), __float_as_int(q1.w), x, y, t, tests, kCount);
            }
        }
    }

    // Vertical walls by a 2-D DDA from the ray origin through the walk list of the ray's sign
    // combination (back-face culling was done when the lists were built).  The loop carries one PENDING
    // record, already loaded; one iteration tests it and fetches the next one: the next record of the
    // current cell, or - when the cell's list is exhausted - the head of the next cell the DDA steps to
    // (first record + continuation in one round trip), all in straight-line predicated code.  In the
    // first version of this loop "test" and "advance" were the two sides of a branch and ran one after
    // the other with 4.5 and 10 of 32 lanes (profiles/r1_bench_synth4000_ncu_summary.csv).  The walk ends
    // when the next cell starts beyond the best hit, where the ray leaves the grid, or where it leaves the
    // z range of the walls (t_exit replaces per-step bounds checks).
    template <bool kCount>
    __device__ __forceinline__ void walk(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                         float dz, unsigned &tests)
    {
        const GridDesc &g = p.grid;
        const float inf = __int_as_float(0x7f800000);
        const bool xp = dx > 0.0f, yp = dy > 0.0f;
        const bool x0 = dx == 0.0f, y0 = dy == 0.0f;
        const float ix = rcp_fast(dx), iy = rcp_fast(dy), iz = rcp_fast(dz);
        int cx = __float2int_rd(fmaf(ox, g.inv_cell, g.bx)), cy = __float2int_rd(fmaf(oy, g.inv_cell, g.by));
        cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1);
        // per-axis DDA constants; an axis the ray does not move along never triggers a step (tmx = inf),
        // so its per-cell increment cell * |1/d| is never used
        float tmx = (fmaf((float)(cx + (xp ? 1 : 0)), g.cell, g.x0) - ox) * ix;
        float tmy = (fmaf((float)(cy + (yp ? 1 : 0)), g.cell, g.y0) - oy) * iy;
        const float ex = ((xp ? g.exit_hi_x : g.exit_lo_x) - ox) * ix, ey = ((yp ? g.exit_hi_y : g.exit_lo_y) - oy) * iy;
        // no wall beyond the z range of the walls (with a little slack for the approximate reciprocal)
        const float ez = ((dz < 0.0f ? g.wall_z_lo : g.wall_z_hi) - oz) * iz * 1.0001f;
        // A ray that starts outside the exit box or the z slab and moves away from it has a negative t_exit
        // (a huge number for the unsigned "0 <= t < best" compare): clamped to 0 nothing passes the test, and
        // with the DDA times clamped alike "the next cell starts beyond best" ends the walk after the first head
        // instead of stepping off the grid.
        tmx = x0 ? inf : fmaxf(tmx, 0.0f); tmy = y0 ? inf : fmaxf(tmy, 0.0f);
        const float t_exit = fmaxf(fminf(fminf(x0 ? inf : ex, y0 ? inf : ey), dz == 0.0f ? inf : ez), 0.0f);
        const int sx = xp ? 1 : -1, sy = yp ? g.nx : -g.nx;
        // current cell as an index into T: the head of the walk list of the ray's sign combination
        int ci = g.walk_base + ((xp ? 1 : 0) + (yp ? 2 : 0)) * g.ncell + cy * g.nx + cx;
        // a ray with d.x == 0 faces no wall whose normal is along x: NaN never passes t < best
        const float nanv = __int_as_float(0x7fc00000);
        const float ax = x0 ? nanv : ix, ay = y0 ? nanv : iy;
        const float bx = -ox * ax, by = -oy * ay;
        // pending record: the head of the origin's cell
        if (!FMGI_CHECK(p, (unsigned)ci < p.grid_records, 4)) ci = g.walk_base;
        float4 q0, h1;
        ldg256(p.grid_table + 2 * ci, q0, h1);
        float qc = h1.x;
        unsigned qtag = __float_as_uint(h1.y);
        int cur = ci, r = __float_as_int(h1.z), rend = __float_as_int(h1.w);
        // One bound for hits and for the walk: every wall lies at least a cell inside the exit box and inside
        // the z slab (with slack), so no hit has t >= t_exit, and "the next cell starts beyond the best hit"
        // also covers "the ray has left the grid".  A walk that ends without a hit restores +inf below.
        const float best_in = best;
        best = fminf(best, t_exit);
        // Written in PTX so that it stays one predicated instruction stream: from the C form nvcc rebuilt
        // nested loops (inner loop over the records of a cell, reconvergence behind it).
        //   test the pending record (a dummy head has c = NaN and fails t < best); best / win by select
        //   adv  = the cell's list is exhausted -> step the DDA unless the next cell starts beyond best,
        //          which ends the walk (best starts at t_exit: no wall lies beyond it)
        //   fetch T[adv ? next cell's head : r] as the new pending record; its last two words are the
        //   new (r, rend): every record carries the range of what follows it in its list
        // MISC (scenes with misc records only): such a record is skipped by the fast test and leaves the
        // loop with misc = its index.
        // The ALU pipe (compares, selects, min/max, integer adds) and the FMA pipe each take one warp
        // instruction every other cycle; the straightforward loop has 22 ALU-pipe instructions of 34 and is
        // bound by that pipe, not by issue.  The balanced form moves seven of them to the FMA pipe: two of the
        // four axis selects become a move plus a predicated add of -0 (x + -0 = x), best / win are updated by
        // predicated moves, the cell index by two predicated integer adds, and the DDA increments read |1/d| of
        // their own axis instead of a selected one.  (p.one = 1 from the constant bank keeps ptxas from folding
        // x * 1 + y back into an ALU-pipe add or select.)
#ifndef FMGI_WALK_BALANCED
#define FMGI_WALK_BALANCED 1
#endif
#if FMGI_WALK_BALANCED
#define FMGI_WALK_SEL_DO                                                                                             \
            "add.rn.f32 dh, %27, 0f80000000;\n\t"                                                                    \
            "@ky add.rn.f32 dh, %26, 0f80000000;\n\t"                                                                \
            "add.rn.f32 oh, %24, 0f80000000;\n\t"                                                                    \
            "@ky add.rn.f32 oh, %23, 0f80000000;\n\t"
#define FMGI_WALK_UPDATE                                                                                             \
            "@ok add.rn.f32 %0, t, 0f80000000;\n\t"                                                                  \
            "@ok mad.lo.s32 %1, %14, %17, 0;\n\t"
#define FMGI_WALK_STEP                                                                                               \
            "@gx mad.lo.s32 %4, %31, %17, %4;\n\t"                                                                   \
            "@gy mad.lo.s32 %4, %32, %17, %4;\n\t"                                                                   \
            "abs.f32 sa, %19;\n\t"                                                                                   \
            "@gx fma.rn.f32 %5, sa, %29, %5;\n\t"                                                                    \
            "abs.f32 sa, %20;\n\t"                                                                                   \
            "@gy fma.rn.f32 %6, sa, %29, %6;\n\t"
#else
#define FMGI_WALK_SEL_DO                                                                                             \
            "selp.f32 dh, %26, %27, ky;\n\t"                                                                         \
            "selp.f32 oh, %23, %24, ky;\n\t"
#define FMGI_WALK_UPDATE                                                                                             \
            "selp.f32 %0, t, %0, ok;\n\t"                                                                            \
            "selp.b32 %1, %14, %1, ok;\n\t"
#define FMGI_WALK_STEP                                                                                               \
            "selp.b32 st, %31, %32, stepx;\n\t"                                                                      \
            "selp.f32 sa, %19, %20, stepx;\n\t"                                                                      \
            "abs.f32 sa, sa;\n\t"                                                                                    \
            "@go add.s32 %4, %4, st;\n\t"                                                                            \
            "@gx fma.rn.f32 %5, sa, %29, %5;\n\t"                                                                    \
            "@gy fma.rn.f32 %6, sa, %29, %6;\n\t"
#endif
        // bounds-checked build: the index of a record that is about to be LOADED (`more`; the index left behind by a
        // walk's last iteration may be the end of T) and lies outside T is replaced by 0, flagged in the misc register
        // (-2) and ends the lane's walk
#ifdef FMGI_CHECKED
#define FMGI_WALK_BOUNDS                                                                                             \
            "setp.ge.u32 oob, %14, %33;\n\t"                                                                         \
            "and.pred oob, oob, more;\n\t"                                                                           \
            "@oob mov.b32 %14, 0;\n\t"                                                                               \
            "@oob mov.b32 %16, -2;\n\t"                                                                              \
            "and.pred more, more, !oob;\n\t"
#define FMGI_WALK_BOUNDS_OPERAND , "r"(p.grid_records)
#else
#define FMGI_WALK_BOUNDS
#define FMGI_WALK_BOUNDS_OPERAND
#endif
#define FMGI_WALK_LOOP(MISC_TEST, MISC_EXIT, COUNT)                                                                       \
        asm volatile(                                                                                                \
            "{\n\t"                                                                                                  \
            ".reg .pred ky, ok, adv, cont, stepx, go, gx, gy, pm, real, more, oob;\n\t"                              \
            ".reg .f32 ak, bk, dh, oh, t, pi, pj, tn, lim, sa;\n\t"                                                  \
            ".reg .b32 tb, bb, st;\n\t"                                                                              \
            ".reg .b64 a;\n\t"                                                                                       \
            "mov.s32 %16, -1;\n\t"                                                                                   \
            "WALK:\n\t"                                                                                              \
            "setp.lt.s32 ky, %13, 0;\n\t"                                                                            \
            "selp.f32 ak, %20, %19, ky;\n\t"                                                                         \
            "selp.f32 bk, %22, %21, ky;\n\t"                                                                         \
            FMGI_WALK_SEL_DO                                                                                         \
            "fma.rn.f32 t, %12, ak, bk;\n\t"                                                                         \
            "fma.rn.f32 pi, t, dh, oh;\n\t"                                                                          \
            "fma.rn.f32 pj, t, %28, %25;\n\t"                                                                        \
            "sub.rn.f32 pi, pi, %8;\n\t"                                                                             \
            "sub.rn.f32 pj, pj, %10;\n\t"                                                                            \
            "abs.f32 pi, pi;\n\t"                                                                                    \
            "abs.f32 pj, pj;\n\t"                                                                                    \
            "mov.b32 tb, t;\n\t"                                                                                     \
            "mov.b32 bb, %0;\n\t"                                                                                    \
            "setp.lt.u32 ok, tb, bb;\n\t"                                                                            \
            "setp.le.and.f32 ok, pi, %9, ok;\n\t"                                                                    \
            "setp.le.and.f32 ok, pj, %11, ok;\n\t"                                                                   \
            MISC_TEST                                                                                                \
            FMGI_WALK_UPDATE                                                                                         \
            COUNT                                                                                                    \
            "setp.ge.s32 adv, %2, %3;\n\t"                                                                           \
            "min.f32 tn, %5, %6;\n\t"                                                                                \
            "setp.lt.f32 cont, tn, %0;\n\t"                                                                          \
            "setp.lt.f32 stepx, %5, %6;\n\t"                                                                         \
            "and.pred go, adv, cont;\n\t"                                                                            \
            "and.pred gx, go, stepx;\n\t"                                                                            \
            "and.pred gy, go, !stepx;\n\t"                                                                           \
            "or.pred more, cont, !adv;\n\t"                                                                          \
            FMGI_WALK_STEP                                                                                           \
            "selp.b32 %14, %4, %2, go;\n\t"                                                                          \
            FMGI_WALK_BOUNDS                                                                                         \
            "mul.wide.s32 a, %14, 32;\n\t"                                                                           \
            "add.s64 a, a, %18;\n\t"                                                                                 \
            "@more ld.global.nc.v8.b32 {%8, %9, %10, %11, %12, %13, %2, %3}, [a];\n\t"                               \
            "selp.u32 %15, 1, 0, more;\n\t"                                                                          \
            MISC_EXIT                                                                                                \
            "@more bra WALK;\n\t"                                                                                    \
            "DONE:\n\t"                                                                                              \
            "}"                                                                                                      \
            : "+f"(best), "+r"(win), "+r"(r), "+r"(rend), "+r"(ci), "+f"(tmx), "+f"(tmy), "+r"(tests), "+f"(q0.x),   \
              "+f"(q0.y), "+f"(q0.z), "+f"(q0.w), "+f"(qc), "+r"(qtag), "+r"(cur), "=r"(more), "=r"(misc)            \
            : "r"(p.one), "l"(p.grid_table), "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(ox), "f"(oy), "f"(oz), "f"(dx),     \
              "f"(dy), "f"(dz), "f"(g.cell), "f"(t_exit), "r"(sx), "r"(sy) FMGI_WALK_BOUNDS_OPERAND)
        // counting variant only (fmgi_options.count_tests): rectangle tests, dummy heads (c = NaN) excluded
#define FMGI_WALK_COUNT "setp.eq.f32 real, %12, %12;\n\t@real add.u32 %7, %7, 1;\n\t"
#define FMGI_WALK_MISC_TEST "and.b32 st, %13, 0x40000000;\n\tsetp.ne.u32 pm, st, 0;\n\tand.pred ok, ok, !pm;\n\t@pm mov.b32 %16, %14;\n\t"
        int more, misc;
        if (!p.grid_has_misc) {
            if (kCount) FMGI_WALK_LOOP("", "", FMGI_WALK_COUNT); else FMGI_WALK_LOOP("", "", "");
        } else {
            do {
                if (kCount) FMGI_WALK_LOOP(FMGI_WALK_MISC_TEST, "@pm bra DONE;\n\t", FMGI_WALK_COUNT);
                else FMGI_WALK_LOOP(FMGI_WALK_MISC_TEST, "@pm bra DONE;\n\t", "");
                if (misc >= 0) {
                    // The fast test skipped misc record `misc`; the next pending record is already loaded.
                    // The DDA decided with the old best: at worst it visits one cell more than needed.
                    const float4 m0 = __ldg(p.grid_table + 2 * misc);
                    const float2 m1 = ldg2(p.grid_table + 2 * misc + 1);
                    const float t = grid_test_misc(p.general, m0, m1.x, __float_as_uint(m1.y), ox, oy, oz, dx, dy, dz, best);
                    if (t < best) { best = t; win = misc; }
                }
            } while (misc >= 0 && more);
        }
#ifdef FMGI_CHECKED
        if (misc == -2) fmgi_report(p, 5);
#endif
#undef FMGI_WALK_BOUNDS
#undef FMGI_WALK_BOUNDS_OPERAND
#undef FMGI_WALK_COUNT
#undef FMGI_WALK_MISC_TEST
#undef FMGI_WALK_SEL_DO
#undef FMGI_WALK_UPDATE
#undef FMGI_WALK_STEP
        if (win < 0) best = best_in;
#undef FMGI_WALK_LOOP
    }

    // Wall index of the winner (-1: miss) and its distance recomputed with the reference's formula.
    __device__ __forceinline__ int finish(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                          float dz, float &t_out) const
    {
        int id = -1;
        t_out = best;
        if (win >= 0 && FMGI_CHECK(p, (unsigned)win < p.grid_records, 6)) {
            const float2 q1 = ldg2(p.grid_table + 2 * win + 1);
            const unsigned tag = __float_as_uint(q1.y);
            if ((tag & (kTagMisc | kTagHorizontal)) == (kTagMisc | kTagHorizontal)) {
                const float4 *gg = p.general + 4 * (tag & kTagIdMask);
                const float4 g0 = __ldg(gg), g3 = __ldg(gg + 3);
                id = __float_as_int(g3.w);
                const float denom = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, dx), __fmul_rn(g0.y, dy)), __fmul_rn(g0.z, dz));
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, __fsub_rn(g3.x, ox)),
                                                      __fmul_rn(g0.y, __fsub_rn(g3.y, oy))),
                                            __fmul_rn(g0.z, __fsub_rn(g3.z, oz)));
                t_out = __fdiv_rn(num, denom);
            } else {
                id = (int)(tag & kTagIdMask);
                const bool kz = (tag & (kTagMisc | kTagHorizontal)) != 0, ky = (int)tag < 0;
                const float ok = kz ? oz : (ky ? oy : ox);
                const float dk = kz ? dz : (ky ? dy : dx);
                t_out = __fdiv_rn(__fsub_rn(q1.x, ok), dk);
            }
        }
        return id;
    }
};

template <bool kCount>
__device__ __forceinline__ int closest_hit_grid(const TraceParams &p, float ox, float oy, float oz,
                                                float dx, float dy, float dz, float &t_out, unsigned &tests)
{
    // walls first: the walk is bounded by the z range of the walls, and only the planes crossed before
    // the wall hit need a lookup
    GridWalk w;
    w.reset();
    w.template walk<kCount>(p, ox, oy, oz, dx, dy, dz, tests);
    w.template planes<kCount>(p, ox, oy, oz, dx, dy, dz, tests);
    return w.finish(p, ox, oy, oz, dx, dy, dz, t_out);
}

// Soup tier with the grid's plane tables: horizontal rectangles (floors, ceilings, sills: about a third
// of a flat's soup) are found by one cell lookup per z plane exactly as in the grid tier, which also
// bounds `best` before the shared-memory scan of the vertical walls (x and y lists only).  Used when
// every horizontal rectangle fits the plane table.
template <bool kCount>
__device__ __forceinline__ int closest_hit_soup_planes(const SoupTables &s, const TraceParams &p, float ox, float oy,
                                                       float oz, float dx, float dy, float dz, float &t_out,
                                                       unsigned &tests)
{
    GridWalk w;
    w.reset();
    w.template planes<kCount>(p, ox, oy, oz, dx, dy, dz, tests);
    const float nanv = __int_as_float(0x7fc00000);
    float best = w.best;
    int code = -1;
    const float ax = dx != 0.0f ? __frcp_rn(dx) : nanv;
    const float ay = dy != 0.0f ? __frcp_rn(dy) : nanv;
    const int sx = dx > 0.0f ? 3 : 0, sy = dy > 0.0f ? 3 : 0;
    scan_axis(s.axis + sx, s.pair_begin[0], s.pair_begin[1], ax, -ox * ax, oy, dy, oz, dz, best, code);
    scan_axis(s.axis + sy, s.pair_begin[1], s.pair_begin[2], ay, -oy * ay, ox, dx, oz, dz, best, code);
    if (s.num_general)
        scan_general(s.general, s.num_general, ox, oy, oz, dx, dy, dz, best, code);
    if (code != -1) return soup_winner(s, code, best, ox, oy, oz, dx, dy, dz, t_out);
    t_out = best;
    if (w.win < 0) return -1;
    const float2 q1 = ldg2(p.grid_table + 2 * w.win + 1);
    t_out = __fdiv_rn(__fsub_rn(q1.x, oz), dz);
    return (int)(__float_as_uint(q1.y) & kTagIdMask);
}

// ---- closest hit through the box decomposition (room tier, room_tables.h) -----------------------------------------

constexpr int kRoomWalking = -2;          // rooms_walk: the ray is in box `box`, still on its way
constexpr int kRoomMaxSteps = 4096;       // boxes per ray before the photon is given up (never reached by a sane table)
constexpr unsigned kRoomKindShift = 30;   // room_tables.h: kRoomCode*

// Box a ray that starts at (x, y, z) and travels along d is in: kd-tree descent; a point exactly on a split plane
// belongs to the side the ray travels towards.  The point is clamped into the root box first.
__device__ __forceinline__ int rooms_locate(const TraceParams &p, float x, float y, float z, float dx, float dy, float dz)
{
    x = fminf(fmaxf(x, p.room_lo[0]), p.room_hi[0]);
    y = fminf(fmaxf(y, p.room_lo[1]), p.room_hi[1]);
    z = fminf(fmaxf(z, p.room_lo[2]), p.room_hi[2]);
    int n = 0;
    for (int guard = 0; guard < 256; guard++) {
        if (!FMGI_CHECK(p, (unsigned)n < p.room_num_nodes, 20)) return 0;
        const float4 nd = __ldg(p.room_nodes + n);
        const int axis = __float_as_int(nd.y);
        if (axis < 0) return __float_as_int(nd.z);
        const float c = axis == 0 ? x : (axis == 1 ? y : z);
        const float dc = axis == 0 ? dx : (axis == 1 ? dy : dz);
        const bool right = c > nd.x || (c == nd.x && dc > 0.0f);
        n = right ? __float_as_int(nd.w) : __float_as_int(nd.z);
    }
    return 0;
}

// What is at (pu, pv) on a face (or an emitter rectangle) that holds several things: the cell of grid record `g` the
// point falls into.  {su0, su1, su2, sv0}, {sv1, sv2, base, stride}; unused splits are +inf.
__device__ __forceinline__ unsigned rooms_grid_lookup(const TraceParams &p, const float4 *__restrict__ grids,
                                                      const unsigned *__restrict__ cells, unsigned g, float pu, float pv)
{
    if (!FMGI_CHECK(p, g < p.room_num_face_grids, 22)) return 3u << 30;
    float4 g0, g1;
    ldg256(grids + 2 * (size_t)g, g0, g1);
    // cell = base + (u splits the coordinate is at or above) + stride * (v splits ...): six compares, each followed by
    // a predicated add (from C the compiler builds 0 / 1 values and sums them: 28 instead of 18 instructions a lookup)
    unsigned cell = __float_as_uint(g1.z);
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.f32 p, %1, %3;\n\t@p add.u32 %0, %0, 1;\n\t"
        "setp.ge.f32 p, %1, %4;\n\t@p add.u32 %0, %0, 1;\n\t"
        "setp.ge.f32 p, %1, %5;\n\t@p add.u32 %0, %0, 1;\n\t"
        "setp.ge.f32 p, %2, %6;\n\t@p add.u32 %0, %0, %9;\n\t"
        "setp.ge.f32 p, %2, %7;\n\t@p add.u32 %0, %0, %9;\n\t"
        "setp.ge.f32 p, %2, %8;\n\t@p add.u32 %0, %0, %9;\n\t"
        "}"
        : "+r"(cell)
        : "f"(pu), "f"(pv), "f"(g0.x), "f"(g0.y), "f"(g0.z), "f"(g0.w), "f"(g1.x), "f"(g1.y), "r"(__float_as_uint(g1.w)));
    if (!FMGI_CHECK(p, cell < p.room_num_face_cells, 24)) return 3u << 30;
    return __ldg(cells + cell);
}

// The same descent out of line, for the one ray in 1e5 whose walk finds its origin outside the box it was taken to be
// in (rooms_walk): kept out of the walk's instruction stream.
__device__ __noinline__ int rooms_relocate(const float4 *__restrict__ nodes, float lox, float loy, float loz, float hix, float hiy,
                                           float hiz, float x, float y, float z, float dx, float dy, float dz)
{
    x = fminf(fmaxf(x, lox), hix);
    y = fminf(fmaxf(y, loy), hiy);
    z = fminf(fmaxf(z, loz), hiz);
    int n = 0;
#pragma unroll 1
    for (int guard = 0; guard < 256; guard++) {
        const float4 nd = __ldg(nodes + n);
        const int axis = __float_as_int(nd.y);
        if (axis < 0) return __float_as_int(nd.z);
        const float c = axis == 0 ? x : (axis == 1 ? y : z);
        const float dc = axis == 0 ? dx : (axis == 1 ? dy : dz);
        const bool right = c > nd.x || (c == nd.x && dc > 0.0f);
        n = right ? __float_as_int(nd.w) : __float_as_int(nd.z);
    }
    return 0;
}

// First box of a photon emitted by `emitter`: the boxes in front of the emitter rectangle partition it - one box (a
// ceiling light, a window in its niche: no lookup), or a grid lookup with the start point's two in-plane
// coordinates; an emitter the builder could not place falls back to the kd-tree of the boxes.
__device__ __forceinline__ int rooms_start(const TraceParams &p, int emitter, float x, float y, float z, float dx, float dy,
                                           float dz)
{
    const int2 st = __ldg(p.room_starts + emitter);
    unsigned code = (unsigned)st.x;
    const float pu = st.y == 0 ? y : x, pv = st.y == 2 ? y : z;
    while ((code >> kRoomKindShift) == 0u) code = rooms_grid_lookup(p, p.room_face_grids, p.room_face_cells, code, pu, pv);
    if ((code >> kRoomKindShift) == 2u) return (int)(code & ((1u << kRoomKindShift) - 1u));
    return rooms_locate(p, x, y, z, dx, dy, dz);
}

// 1 / d for the slab tests of the walk.  A ray parallel to an axis never reaches that axis' faces: d = 0 takes a
// huge NEGATIVE finite value - the slab test then reads the box's lower bound (d > 0 is false), lower bound minus
// origin is <= 0, and the product is a huge positive ray parameter (or 0 for an origin exactly on that face, which
// hands the ray to the neighbour it touches).
__device__ __forceinline__ float rooms_inv(float d) { return d == 0.0f ? -1e30f : rcp_fast(d); }


// The ray leaves its box through the nearest of the three faces it travels towards; the face's code - after a
// lookup in the face's grid record where several things share the face - says what is at the exit point: a
// collider that faces into the box (a hit - the closest one: nothing lies inside a box), the next box, or nothing
// (the ray leaves the scene).  At most kSteps boxes per call: returns the wall index, -1 (miss), or kRoomWalking
// with `box` = the box the ray is in now, so that the caller can interleave the walks of a warp's lanes with their
// shading instead of waiting for the longest walk (k_trace).  On a hit `box` is the box the hit was found in: the
// bounce starts there.  (ix, iy, iz) = rooms_inv of the direction.
// A collider whose plane lies BEHIND the origin (negative ray parameter) is not a hit (rectangle.c:76 rejects t < 0):
// the origin is not in the box the walk took it to be in, and the walk goes on from the box it is in (kd-tree descent).
template <bool kCount, int kSteps>
__device__ __forceinline__ int rooms_walk(const TraceParams &p, int &box, float ox, float oy, float oz, float dx, float dy,
                                          float dz, float ix, float iy, float iz, float &t_out, unsigned &tests)
{
    // the ray's octant picks the 32-byte record of each box: far planes and face codes of the three faces ahead
    const float4 *base = p.room_boxes + ((dx > 0.0f ? 2 : 0) + (dy > 0.0f ? 4 : 0) + (dz > 0.0f ? 8 : 0));
    const float4 *face_grids = p.room_face_grids;
    const unsigned *face_cells = p.room_face_cells;
    unsigned cur = (unsigned)box;
    int s = 0;
    unsigned code;
    float t, ca, oa, da;                                 // exit face of the last box: ray parameter, plane, origin / direction along its axis
#pragma unroll 1
    for (;;) {
        if (!FMGI_CHECK(p, cur < p.room_num_boxes, 21)) { code = 3u << kRoomKindShift; break; }
        float4 r0, r1;                                   // {far.x, far.y, far.z, code.x}, {code.y, code.z, -, -}
        ldg256(base + (size_t)kRoomBoxVec * cur, r0, r1);
        if (kCount) tests++;                             // counted: boxes crossed + face grids looked up
        const float tx = (r0.x - ox) * ix, ty = (r0.y - oy) * iy, tz = (r0.z - oz) * iz;
        // nearest face; the two in-plane coordinates of the exit point, in ascending axis order
        const float txy = fminf(tx, ty);
        const bool ax_y = ty < tx, ax_z = tz < txy;
        t = fminf(txy, tz);
        const float hx = fmaf(t, dx, ox), hy = fmaf(t, dy, oy), hz = fmaf(t, dz, oz);
        const float pu = (ax_y || ax_z) ? hx : hy, pv = ax_z ? hy : hz;
        code = __float_as_uint(ax_z ? r1.y : (ax_y ? r1.x : r0.w));
        ca = ax_z ? r0.z : (ax_y ? r0.y : r0.x);
        oa = ax_z ? oz : (ax_y ? oy : ox);
        da = ax_z ? dz : (ax_y ? dy : dx);
        while ((code >> kRoomKindShift) == 0u) {         // several things on this face: look the exit point up
            if (kCount) tests++;
            code = rooms_grid_lookup(p, face_grids, face_cells, code, pu, pv);
        }
        if ((code >> kRoomKindShift) != 2u) break;       // a collider or nothing: the walk ends here
        cur = code & ((1u << kRoomKindShift) - 1u);
        if (++s == kSteps) { box = (int)cur; return kRoomWalking; }
    }
    // all lanes whose walk ended, together: a hit needs the collider in front of the origin
    box = (int)cur;
    if ((code >> kRoomKindShift) != 1u) return -1;       // nothing there: the ray leaves the scene
    if (!(t >= 0.0f)) {
        // The wall's plane lies BEHIND the origin: the origin is not in this box but behind that wall - the 1e-5 step
        // along the ray put a photon beside its window's niche, or rounding put a bounce point an ulp beyond the wall
        // of a corner.  The reference does not hit such a wall (rectangle.c:76) and lets the photon fly on inside the
        // wall's material; so do we: the walk goes on from the box the origin really is in (about one ray in 1e5).
        box = rooms_relocate(p.room_nodes, p.room_lo[0], p.room_lo[1], p.room_lo[2], p.room_hi[0], p.room_hi[1], p.room_hi[2],
                             ox, oy, oz, dx, dy, dz);
        return kRoomWalking;
    }
    // (a ray parallel to the collider it starts on is culled like a back face, rectangle.c:70-72: n.d = 0)
    if (da == 0.0f) return -1;
    // the distance with the reference's formula for an axis-parallel normal, IEEE division
    t_out = __fdiv_rn(__fsub_rn(ca, oa), da);
    return (int)(code & ((1u << kRoomKindShift) - 1u));
}

// The whole walk of one ray (probes, ambient occlusion).
template <bool kCount = false>
__device__ __forceinline__ int closest_hit_rooms(const TraceParams &p, int &leaf, float ox, float oy, float oz, float dx,
                                                 float dy, float dz, float &t_out, unsigned *tests = nullptr)
{
    unsigned n = 0;
    t_out = __int_as_float(0x7f800000);
    const float ix = rooms_inv(dx), iy = rooms_inv(dy), iz = rooms_inv(dz);
    int r = kRoomWalking;
    // (a walk comes back "still walking" when it had to find the origin's box again; twice is already unheard of)
    for (int again = 0; again < 4 && r == kRoomWalking; again++)
        r = rooms_walk<kCount, kRoomMaxSteps>(p, leaf, ox, oy, oz, dx, dy, dz, ix, iy, iz, t_out, n);
    if (kCount && tests) *tests += n;
    return r == kRoomWalking ? -1 : r;
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

__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Malley disk sample around n with the precomputed basis (u, v).  The phi product rounds like the reference's
// double-then-float evaluation; sin / cos and the two square roots use the SFU (sin / cos are good to 2^-21, so a
// correctly rounded square root - seven more instructions each - bought no closer agreement with the reference's
// directions: identical 8-bounce paths 99.96 % either way).
__device__ __forceinline__ void sample_hemisphere(float xi1, float xi2, bool sky, const float4 n, const float4 u,
                                                  const float4 v, float &dx, float &dy, float &dz)
{
    const float r = sqrt_fast(xi1);
    const float phi = __fmul_rn(2.0f * 3.141592f, xi2);
    float sp, cp;
    __sincosf(phi, &sp, &cp);
    float a = r * cp;
    const float b = r * sp;
    const float c = sqrt_fast(__fsub_rn(1.0f, __fmul_rn(r, r)));
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
