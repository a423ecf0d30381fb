// Device-side scene tables derived from the reference's Rectangle soup (rectangle.h:19-26).
// Built once per scene on the host (scene_prep.cpp), uploaded, then staged into shared memory by
// the trace kernel (soup tier) or walked from L2 (grid tier).
#pragma once
#include <stdint.h>
#include <vector>

#include "../../include/fmgi.h"

namespace fmgi {

// ---- closest-hit tables -------------------------------------------------------------------

// Colliders whose width/height/normal are axis parallel (everything parseLayout.c emits) are
// stored per normal axis k as two lists: P (normal +k, hit only by rays with d[k] < 0) and
// M (normal -k, d[k] > 0) - back-face culling (rectangle.c:70-72) becomes "each lane walks the
// list its ray can face".  The lists are padded to the same even length and interleaved in
// blocks of two rectangles, so that a warp executes one uniform loop while every lane reads the
// block of its own sign: block(k, j, s) = blocks[(pair_begin[k] + j) * 2 + s].
// (i, j) are the two in-plane axes in ascending order; extents are kept as centre and
// half-width so that containment is |p - mid| <= half (edges inclusive, rectangle.c:93).
// 48 bytes = three 16-byte shared-memory loads per two tests; the P and M block of one j are
// 48 bytes apart, i.e. in disjoint banks.
struct AxisPairBlock {
    float c_a, mid_i_a, half_i_a, mid_j_a;    // rectangle a: plane coordinate pos[k], extents
    float c_b, mid_i_b, half_i_b, mid_j_b;    // rectangle b
    float half_j_a, half_j_b;
    int32_t id_a, id_b;                       // index into the caller's wall table (-1: padding)
};
static_assert(sizeof(AxisPairBlock) == 48, "AxisPairBlock is three float4");

// Arbitrarily oriented collider: the full plane + two projections of rectangle.c:67-95.
struct GeneralRect {
    float nx, ny, nz, nd;       // unit normal, n . pos
    float wx, wy, wz, wlen;     // width / |width|, |width|
    float hx, hy, hz, hlen;     // height / |height|, |height|
    float px, py, pz;           // pos
    int32_t id;
};

// ---- shading tables (read once per hit / per emission) ---------------------------------------

// Everything the bounce needs about the wall that was hit, as six float4 (96 B).  q0..q2 feed the
// texel index, q3..q5 the sampler frame; EmitterRec shares the q3..q5 layout so that emission
// and re-emission run through the same code.
struct ShadeRect {
    float pos[3]; int32_t base;      // q0: pos, atlas base index lightmapSetup[0]
    float wn[3];  float wlen;        // q1: width * (1/|width|) as div_vec3 computes it, |width|
    float hn[3];  float hlen;        // q2: height * (1/|height|), |height|
    float n[3];   int32_t tiles;     // q3: unit normal as stored by the reference; tiles_w | tiles_h << 16
    float u[3];   int32_t pad0;      // q4: sampler basis U (vector3_cl.c:139-144)
    float v[3];   int32_t pad1;      // q5: sampler basis V
};
static_assert(sizeof(ShadeRect) == 96, "ShadeRect is six float4");

struct EmitterRec {
    float pos[3];    int32_t is_window;   // q0  photonmap.c:169-171,179-181
    float width[3];  float pad0;          // q1
    float height[3]; float pad1;          // q2
    float n[3];      float pad2;          // q3
    float u[3];      float pad3;          // q4
    float v[3];      float pad4;          // q5
};
static_assert(sizeof(EmitterRec) == 96, "EmitterRec is six float4");

// ---- uniform grid over the floor plan (grid tier) --------------------------------------------
//
// parseLayout.c extrudes a 2-D floor plan, so a 2-D grid over (x, y) is the natural index:
//   * horizontal rectangles (floors, ceilings, sills, lintels) live on a handful of z planes.
//     For every distinct (z, normal sign) plane there is one cell list per grid cell; a ray finds
//     its crossing point with the plane, looks up that one cell and tests only its candidates;
//   * everything else (vertical walls, arbitrarily oriented rectangles, horizontal rectangles
//     beyond the plane table) is binned by its (x, y) bounding box into the "walk" lists of every
//     cell it overlaps and found by a 2-D DDA from the ray origin, which stops as soon as the
//     next cell starts beyond the best hit so far.  There are four walk lists per cell, one per
//     sign combination (d.x > 0, d.y > 0) of the ray: a vertical wall is stored only in the two
//     lists whose rays can face it (back-face culling, rectangle.c:70-72, done at build time).
// One table T of 32-byte records serves every list: entry l * ncell + cell is the HEAD of list
// (l, cell) - its first record stored inline (or a dummy that no ray can hit) together with the index
// range [next, end) of the list's remaining records, which follow the heads in T; every record carries
// the range of what follows IT, so that fetching a record also fetches the walk's continuation state.  A cell visit is
// therefore ONE dependent memory round trip (head = first candidate + continuation) instead of two
// (range, then record).  Lists are numbered compactly: [0, planes_up) planes with normal +z (highest
// first), [planes_up, planes_up + planes_down) planes with normal -z (lowest first), then the four
// walk lists; GridDesc carries the bases.  (build_grid also keeps the CSR form it builds T from:
// grid_ranges / grid_recs, lists numbered l < 16: planes, l = 16 + combo: walk lists.)

enum { kMaxPlanesPerSign = 8, kWalkListBase = 2 * kMaxPlanesPerSign, kNumGridLists = kWalkListBase + 4 };

// Record tag: wall id in the low 28 bits plus
//   bit 31       vertical wall whose normal is along y (clear: along x)  - the sign bit, one compare
//   bit 30       "misc" record that takes the slow path of the walk: bit 29 set = arbitrarily oriented
//                rectangle (low bits index `general`), bit 29 clear = horizontal rectangle beyond the
//                plane table (bit 28: normal is -z)
//   bit 29 alone horizontal rectangle in a plane list
enum : uint32_t { kTagAlongY = 1u << 31, kTagMisc = 1u << 30, kTagHorizontal = 1u << 29, kTagNegative = 1u << 28,
                  kTagIdMask = (1u << 28) - 1 };

// The first float4 is all a plane lookup needs (extents); the walk also reads (c, tag).
struct GridRec {
    float mid_i, half_i;    // extent along in-plane axis i (centre, half width)
    float mid_j, half_j;
    float c;                // plane coordinate pos[k]; NaN in a dummy head
    uint32_t tag;
    int32_t next, end;      // the records that follow this one in its list: T[next .. end) (next == end: none)
};
static_assert(sizeof(GridRec) == 32, "GridRec is two float4");

struct GridDesc {
    float x0, y0;           // world position of cell (0,0)'s corner
    float cell, inv_cell;   // cell edge, 1/edge
    int32_t nx, ny;
    int32_t planes_up, planes_down;         // number of z planes with normal +z / -z
    float plane_z[2 * kMaxPlanesPerSign];   // [0, planes_up): normal +z; [kMaxPlanesPerSign, +planes_down): normal -z
    // derived constants of the walk
    int32_t ncell;          // nx * ny
    int32_t planes_max;     // max(planes_up, planes_down)
    float bx, by;           // -x0 * inv_cell, -y0 * inv_cell: cell coordinate = fma(x, inv_cell, bx)
    float exit_lo_x, exit_hi_x, exit_lo_y, exit_hi_y;   // the box without the outermost ring of cells
    float wall_z_lo, wall_z_hi;                         // z range of everything in the walk lists
    int32_t down_base;      // T index of the first head of the first plane with normal -z (planes_up * ncell)
    int32_t walk_base;      // T index of the first head of walk list 0 ((planes_up + planes_down) * ncell)
    // The first three planes a ray travelling down (index 0: faces normals +z) / up (index 1) can meet, nearest
    // first, as the plane lookup reads them with one indexed constant load each: plane coordinates (NaN in an
    // unused slot: its ray parameter is NaN and fails every compare) and T indices of their first heads.
    float fast_z[2][4];
    int32_t fast_base[2][4];
};

struct HostScene {
    int num_walls = 0, num_windows = 0, num_lights = 0, num_texels = 0;
    std::vector<AxisPairBlock> axis;      // 2 blocks (P, M) per pair; pairs of axis k: pair_begin[k] .. pair_begin[k+1]
    int pair_begin[4] = {0, 0, 0, 0};
    int num_axis_rects = 0;               // real (unpadded) axis-parallel colliders
    std::vector<GeneralRect> general;
    std::vector<ShadeRect> shade;         // per wall
    std::vector<EmitterRec> emitters;     // windows then lights
    std::vector<float> emitter_area;      // |w|*|h| in float (photonmap.c:417)
    // grid tier (filled by build_grid)
    GridDesc grid = {};
    std::vector<int32_t> grid_ranges;     // 2 ints (begin, end) per (list, cell); kNumGridLists lists
    std::vector<GridRec> grid_recs;
    std::vector<GridRec> grid_table;      // T: what the device walks
    int grid_overflow_horizontal = 0;     // horizontal rectangles that did not fit the plane table
    int grid_misc = 0;                    // misc records in the walk lists (those plus arbitrarily oriented rectangles)
};

// photonmap.c:414-418: N = (uint64)(int spa * float area)
uint64_t photon_budget(float area, int samples_per_area);

// Returns an empty string on success, else the reason the scene was rejected.
const char *prepare_scene(HostScene &out, const fmgi_rect *walls, int num_walls,
                          const fmgi_rect *windows, int num_windows,
                          const fmgi_rect *lights, int num_lights, int num_texels);

// One collider as build_grid bins it: list < 0: the walk lists its rays can face (axis 0 / 1: the two sign
// combinations that face normal sign `neg`; axis 2 or 3: all four), else the plane list; cells [cx0, cx1] x [cy0, cy1].
struct GridItem {
    int32_t list;
    int32_t axis, neg;
    int32_t cx0, cx1, cy0, cy1;
    int32_t pad;
    GridRec rec;
};
static_assert(sizeof(GridItem) == 64, "GridItem is 64 bytes");

// Builds the floor-plan grid of the grid tier.  cell_hint <= 0 picks the cell edge from the scene
// (about two colliders per cell).  Needs the wall table again because HostScene keeps only
// derived records.
void build_grid(HostScene &scene, const fmgi_rect *walls, int num_walls,
                const fmgi_rect *windows, int num_windows, const fmgi_rect *lights, int num_lights,
                float cell_hint);

// The two halves of build_grid (scene_prep.cpp): per-collider classification + GridDesc, and the per-cell assembly
// of T, which also exists as device kernels (grid_build.cuh) for large scenes.
void grid_classify(HostScene &scene, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows, int num_windows,
                   const fmgi_rect *lights, int num_lights, float cell_hint, std::vector<GridItem> &items);
void grid_assemble_host(HostScene &scene, const std::vector<GridItem> &items);

// Geodesic half-sphere direction set (xyz triples); iterations = 4 is the reference's geoSphere4.
std::vector<float> geosphere_directions(int iterations);

// vector3_cl.c:139-144: the basis both hemisphere samplers build around a normal.
void sampler_basis(const float n[3], float u[3], float v[3]);

}  // namespace fmgi
