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
//   * the rectangle soup lives in shared memory as per-axis lists of axis-parallel records; the
//     loop trip count is warp uniform and the loads are broadcasts;
//   * back-face culling (rectangle.c:70-72) halves the work instead of costing a test: the
//     records of one normal axis are split by normal sign into two interleaved lists and every
//     lane walks only the list its ray can face (two distinct, bank-disjoint addresses per warp);
//   * 0 <= t < best is one unsigned compare (negative and NaN floats are large unsigned
//     integers); containment is |p - mid| <= half, which moves half of the compare work from the
//     ALU pipe to the FMA pipe (the ALU pipe was the top pipe of the first version, see profiles/);
//   * horizontal rectangles (a third of a flat) are not scanned at all when they fit the grid's
//     plane tables: one cell lookup per z plane at the ray's crossing point (GridWalk::planes);
//   * scenes beyond a few hundred rectangles use the floor-plan grid instead of the soup
//     (GridWalk: plane lookups + a 2-D DDA through per-cell, per-sign-combination wall lists);
//   * deposits are one 16-byte vector reduction (RED.E.ADD.F32x4) per bounce into the atlas,
//     optionally warp-aggregated with __match_any_sync (measured: no gain, not the default);
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
// internal kernel variant: soup tier whose horizontal rectangles go through the grid's plane tables
constexpr int kTierSoupPlanes = 3;

struct TraceParams {
    // closest-hit tables (global memory; staged into shared memory by the soup kernel)
    const float4 *axis;          // 3 float4 per AxisPairBlock, 2 blocks (P, M) per pair
    const float4 *general;       // 4 float4 per GeneralRect
    int pair_begin[4];           // pairs of axis k: pair_begin[k] .. pair_begin[k+1]
    int num_general;
    // grid tier (scene_tables.h): inline cell-list records and (begin, end) per (list, cell)
    const float4 *grid_recs;     // 2 float4 per GridRec
    const int2 *grid_ranges;
    GridDesc grid;
    // shading tables
    const float4 *shade;         // 6 float4 per wall
    const float4 *emitters;      // 6 float4 per emitter
    const float *ao_width;       // ambient occlusion only: the walls' width / height vectors (3 floats each)
    const float *ao_height;
    // this shard's photon index space: jobs [job_begin[e], job_begin[e+1]) belong to emitter e and
    // map to photon indices photon_first[e] + (job - job_begin[e])
    const unsigned long long *job_begin;
    const unsigned long long *photon_first;
    int num_emitters;
    unsigned long long total_jobs;
    unsigned long long *work_counter;
    // output
    float4 *atlas;
    unsigned long long *counters;   // photons, rays, deposits, mirror bounces, (work counter), rectangle tests (grid tier)
    int32_t *path_out;              // probe builds only
    int max_depth;
    uint32_t seed;
};

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

struct GridHit { float best; int rec; };      // rec: index of the winning inline record, -1 = miss

// Rare walk-list records: a horizontal rectangle beyond the plane table (k == 2) or an arbitrarily
// oriented rectangle (k == 3).  These sit in all four walk lists, so facing is tested here.
// Returns the ray parameter of a valid hit closer than `best`, else +inf (by value: the caller's
// running minimum stays in registers).
__device__ __noinline__ float grid_test_misc(const float4 *__restrict__ general, float4 q0, float4 q1, float ox, float oy,
                                             float oz, float dx, float dy, float dz, float best)
{
    const float inf = __int_as_float(0x7f800000);
    const int tag = __float_as_int(q1.y);
    if (((tag >> 28) & 3) == 3) {
        // rectangle.c:67-95 for an arbitrarily oriented rectangle
        const float4 *g = general + 4 * (tag & 0x0fffffff);
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
    const bool facing = (tag & (1 << 30)) ? dz > 0.0f : dz < 0.0f;    // back-face culling, rectangle.c:70-72
    const float t = __fdividef(q0.x - oz, dz);
    const float pi = fmaf(t, dx, ox) - q0.y;
    const float pj = fmaf(t, dy, oy) - q0.w;
    const bool ok = facing && (__float_as_uint(t) < __float_as_uint(best)) && fabsf(pi) <= q0.z && fabsf(pj) <= q1.x;
    return ok ? t : inf;
}

// Closest-hit query through the grid as begin / step / finish.  (A kernel that interleaved the
// steps of different rays - lanes whose walk was over waited until most of the warp was waiting,
// then shaded together - was measured and dropped: 54.3 ms vs 53.7 ms for this one-ray-at-a-time
// form on the 21.5k-rectangle scene; the bookkeeping ate what the fuller warps gained.)
struct GridWalk {
    float best;            // ray parameter of the best hit so far (+inf: none)
    int win;               // inline record index of the best hit, -1: none
    float ix, iy;          // 1/d.x, 1/d.y
    float tmx, tmy;        // ray parameter at which the walk leaves the current cell along x / y
    float tdx, tdy;        // ray parameter per cell along x / y
    float t_exit;          // ray parameter at which the ray leaves the grid minus its outermost ring of cells
    int ci;                // current cell, linear index cy * nx + cx
    int sx, sy;            // linear-index step along x (+-1) and y (+-nx)
    int r, rend;           // pending records of the current cell
    const int2 *walk;      // the walk list ranges of this ray's sign combination

    // Phase 1: horizontal planes the ray can face - one cell lookup per plane at the crossing point.
    __device__ __forceinline__ void planes(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                           float dz, unsigned &tests)
    {
        const GridDesc &g = p.grid;
        const int ncell = g.nx * g.ny;
        best = __int_as_float(0x7f800000);
        win = -1;
        if (dz != 0.0f) {
            const float iz = __frcp_rn(dz);
            const int first = dz < 0.0f ? 0 : kMaxPlanesPerSign;             // d.z < 0 faces normals +z
            const int count = dz < 0.0f ? g.planes_up : g.planes_down;
            for (int pl = 0; pl < count; pl++) {
                const float t = (g.plane_z[first + pl] - oz) * iz;
                if (!(__float_as_uint(t) < __float_as_uint(best))) continue;
                const float x = fmaf(t, dx, ox), y = fmaf(t, dy, oy);
                const int px = __float2int_rd((x - g.x0) * g.inv_cell), py = __float2int_rd((y - g.y0) * g.inv_cell);
                if (px < 0 || py < 0 || px >= g.nx || py >= g.ny) continue;
                const int2 range = __ldg(p.grid_ranges + (first + pl) * ncell + py * g.nx + px);
                for (int q = range.x; q < range.y; q++) {
                    const float4 q0 = __ldg(p.grid_recs + 2 * q);
                    const float4 q1 = __ldg(p.grid_recs + 2 * q + 1);
                    tests++;
                    if (fabsf(x - q0.y) <= q0.z && fabsf(y - q0.w) <= q1.x) { best = t; win = q; break; }
                }
            }
        }
    }

    // Phase 1 + DDA set-up.
    __device__ __forceinline__ void begin(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                          float dz, unsigned &tests)
    {
        const GridDesc &g = p.grid;
        const int ncell = g.nx * g.ny;
        const float inf = __int_as_float(0x7f800000);
        planes(p, ox, oy, oz, dx, dy, dz, tests);
        ix = __frcp_rn(dx); iy = __frcp_rn(dy);
        int cx = __float2int_rd((ox - g.x0) * g.inv_cell), cy = __float2int_rd((oy - g.y0) * g.inv_cell);
        cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1);
        // per-axis DDA constants; an axis the ray does not move along never triggers a step.  t_exit
        // replaces the per-step bounds check: the walk ends where the ray leaves the grid's box.
        tmx = inf; tmy = inf; tdx = inf; tdy = inf; t_exit = inf;
        if (dx != 0.0f) {
            tmx = (g.x0 + (float)(cx + (dx > 0.0f ? 1 : 0)) * g.cell - ox) * ix;
            tdx = g.cell * fabsf(ix);
            t_exit = (g.x0 + (dx > 0.0f ? (float)(g.nx - 1) : 1.0f) * g.cell - ox) * ix;
        }
        if (dy != 0.0f) {
            tmy = (g.y0 + (float)(cy + (dy > 0.0f ? 1 : 0)) * g.cell - oy) * iy;
            tdy = g.cell * fabsf(iy);
            t_exit = fminf(t_exit, (g.y0 + (dy > 0.0f ? (float)(g.ny - 1) : 1.0f) * g.cell - oy) * iy);
        }
        sx = dx > 0.0f ? 1 : -1;
        sy = dy > 0.0f ? g.nx : -g.nx;
        ci = cy * g.nx + cx;
        const int combo = (dx > 0.0f ? 1 : 0) + (dy > 0.0f ? 2 : 0);
        walk = p.grid_ranges + (kWalkListBase + combo) * ncell;
        const int2 range = __ldg(walk + ci);
        r = range.x; rend = range.y;
    }

    // Phase 2, one step of the 2-D DDA through the walk lists of the ray's sign combination: test the
    // next pending record, or move to the next cell.  Cell stepping and record testing are
    // flattened into one loop in which every lane does exactly one thing per step, so lanes with
    // long lists and lanes crossing empty cells keep each other busy (the nested-loop version ran
    // with 4 of 32 lanes active, profiles/r1_v1_grid_ncu_summary.csv).  Returns false when the walk
    // is over: the next cell starts beyond the best hit, or the ray leaves the grid.
    __device__ __forceinline__ bool step(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                         float dz, unsigned &tests)
    {
        if (r < rend) {
            const float4 *rec = p.grid_recs + 2 * r;
            const float4 q0 = __ldg(rec);
            const float4 q1 = __ldg(rec + 1);
            const int tag = __float_as_int(q1.y);
            tests++;
            if (tag & (2 << 28)) {
                const float t = grid_test_misc(p.general, q0, q1, ox, oy, oz, dx, dy, dz, best);
                if (t < best) { best = t; win = r; }
            } else {
                // vertical wall, normal along x (k = 0) or y (k = 1); the list only holds walls this ray
                // can face (back-face culling done at build time).  In-plane axes: the other horizontal
                // axis and z.
                const bool ky = (tag & (1 << 28)) != 0;
                const float t = (q0.x - (ky ? oy : ox)) * (ky ? iy : ix);
                const float pi = fmaf(t, ky ? dx : dy, ky ? ox : oy) - q0.y;
                const float pj = fmaf(t, dz, oz) - q0.w;
                if ((__float_as_uint(t) < __float_as_uint(best)) && fabsf(pi) <= q0.z && fabsf(pj) <= q1.x) {
                    best = t; win = r;
                }
            }
            r++;
            return true;
        }
        const float t_next = fminf(tmx, tmy);
        if (!(t_next < fminf(best, t_exit))) return false;
        if (tmx < tmy) { ci += sx; tmx += tdx; } else { ci += sy; tmy += tdy; }
        const int2 range = __ldg(walk + ci);
        r = range.x; rend = range.y;
        return true;
    }

    // Wall index of the winner (-1: miss) and its distance recomputed with the reference's formula.
    __device__ __forceinline__ int finish(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                          float dz, float &t_out) const
    {
        int id = -1;
        t_out = best;
        if (win >= 0) {
            const float4 q0 = __ldg(p.grid_recs + 2 * win);
            const int tag = __float_as_int(__ldg(p.grid_recs + 2 * win + 1).y);
            const int k = (tag >> 28) & 3;
            if (k == 3) {
                const float4 *gg = p.general + 4 * (tag & 0x0fffffff);
                const float4 g0 = __ldg(gg), g3 = __ldg(gg + 3);
                id = __float_as_int(g3.w);
                const float denom = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, dx), __fmul_rn(g0.y, dy)), __fmul_rn(g0.z, dz));
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(g0.x, __fsub_rn(g3.x, ox)),
                                                      __fmul_rn(g0.y, __fsub_rn(g3.y, oy))),
                                            __fmul_rn(g0.z, __fsub_rn(g3.z, oz)));
                t_out = __fdiv_rn(num, denom);
            } else {
                id = tag & 0x0fffffff;
                const float ok = k == 0 ? ox : (k == 1 ? oy : oz);
                const float dk = k == 0 ? dx : (k == 1 ? dy : dz);
                t_out = __fdiv_rn(__fsub_rn(q0.x, ok), dk);
            }
        }
        return id;
    }
};

__device__ __forceinline__ int closest_hit_grid(const TraceParams &p, float ox, float oy, float oz,
                                                float dx, float dy, float dz, float &t_out, unsigned &tests)
{
    GridWalk w;
    w.begin(p, ox, oy, oz, dx, dy, dz, tests);
    while (w.step(p, ox, oy, oz, dx, dy, dz, tests)) {}
    return w.finish(p, ox, oy, oz, dx, dy, dz, t_out);
}

// Soup tier with the grid's plane tables: horizontal rectangles (floors, ceilings, sills: about a third
// of a flat's soup) are found by one cell lookup per z plane exactly as in the grid tier, which also
// bounds `best` before the shared-memory scan of the vertical walls (x and y lists only).  Used when
// every horizontal rectangle fits the plane table.
__device__ __forceinline__ int closest_hit_soup_planes(const SoupTables &s, const TraceParams &p, float ox, float oy,
                                                       float oz, float dx, float dy, float dz, float &t_out,
                                                       unsigned &tests)
{
    GridWalk w;
    w.planes(p, ox, oy, oz, dx, dy, dz, tests);
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
    const float4 q0 = __ldg(p.grid_recs + 2 * w.win);
    const int tag = __float_as_int(__ldg(p.grid_recs + 2 * w.win + 1).y);
    t_out = __fdiv_rn(__fsub_rn(q0.x, oz), dz);
    return tag & 0x0fffffff;
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
