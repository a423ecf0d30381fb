// Room tier: closest hit by walking a box decomposition of the flat (built by rooms_build.cpp).
//
// Every collider parseLayout.c emits is an axis-parallel rectangle, and a flat is mostly empty boxes: rooms,
// door and window niches, the inside of walls.  A kd-tree splits the scene's bounding box at collider planes
// until NO collider lies inside a leaf box, and leaves that share a whole collider-free face are merged back into
// bigger boxes: all colliders sit on box faces.  A ray then never tests a collider "on spec": it leaves its box
// through one of the three faces it travels towards, and the face says what is at the exit point - a collider that
// faces into the box (hit: the closest one by construction), a neighbouring box (portal: continue there), or
// nothing (the ray leaves the scene).  A ray in a room does ONE step (its own room's walls), a ray through a doorway
// three; no grid walk through empty cells, no per-cell candidate tests, no separate pass over the horizontal planes
// (floors, ceilings, sills and lintels are faces like any other).
//
// What is on a face is a partition of the face into rectangles (colliders, portals, nothing), stored as a small 2-D
// kd-tree over the face's two in-plane coordinates: a face with one thing on it - a floor, a ceiling, a plain wall -
// is a terminal code in the box record itself and costs no lookup at all; a wall with a door is two or three
// compare-and-descend steps.  No containment tests: an exit point that rounding puts an ulp outside the face still
// descends to the nearest part.
// Back-face culling (rectangle.c:70-72) is structural: a face shows only the colliders whose normal points into
// its box.  Where coplanar colliders overlap, the part belongs to the lowest wall index, as the reference's strict
// `<` decides (photonmap.cl:199); a point exactly on the edge between two parts belongs to the part on the higher
// side of the edge.
#pragma once
#include <stdint.h>
#include <vector>

#include "../../include/fmgi.h"

namespace fmgi {

// What a face (or a part of it) holds: kind in the top two bits, index below.
enum : uint32_t {
    kRoomCodeNode = 0u << 30,        // index of a RoomFaceGrid: look the exit point up
    kRoomCodeWall = 1u << 30,        // wall index: a hit
    kRoomCodeBox = 2u << 30,         // box index: the ray goes on there
    kRoomCodeMiss = 3u << 30,        // nothing: the ray leaves the scene
    kRoomCodeKind = 3u << 30,
    kRoomCodeIndex = ~(3u << 30),
};

// One box, as eight 32-byte records: record o serves the rays of octant o = (d.x > 0) | (d.y > 0) << 1 |
// (d.z > 0) << 2 and holds the three faces such a ray can leave through - their plane coordinates and codes - so
// ONE 256-bit load, at an address that depends on the box and the ray's octant only, fetches all a step needs.
struct RoomOctant {
    float far[3];
    uint32_t code[3];
    uint32_t pad[2];
};
struct RoomBox {
    RoomOctant oct[8];
};
static_assert(sizeof(RoomBox) == 256, "RoomBox is sixteen float4");

// One lookup record of a face that holds several things: up to three ascending split values per in-plane coordinate
// (u: the lower in-plane axis, v: the higher; unused splits are +inf) cut the face into up to 4 x 4 cells, and the cell
// the exit point falls into - cell (iu, iv), iu = number of u splits the coordinate is at or above - holds the code of
// what is there: cells[base + iu + stride * iv].  A wall with a door is ONE record (left | door | right) x (opening |
// lintel); a cell that still holds several things (more than three edges along an axis) is the code of a further
// record.  One 256-bit load, six compares, one 32-bit load - no loop over tree levels, whose trip count a warp's
// deepest lane would dictate.
struct RoomFaceGrid {
    float su[3], sv[3];
    uint32_t base, stride;
};
static_assert(sizeof(RoomFaceGrid) == 32, "RoomFaceGrid is two float4");

// Bounds of a box (where does a new photon start; host replay).
struct RoomBounds {
    float lo[3], hi_x;
    float hi_y, hi_z;
    int32_t pad[2];
};
static_assert(sizeof(RoomBounds) == 32, "RoomBounds is two float4");

// Where an emitter's photons start: the boxes in front of the emitter rectangle partition it like the things on a box
// face partition the face - `code` is a box, or a RoomFaceGrid over the rectangle's two in-plane coordinates (normal
// along `axis`), or "nothing" (an emitter that is not axis parallel, or lies outside the
// boxes: the photon's first box is then found by descending the kd-tree below).
struct RoomStart {
    uint32_t code;
    int32_t axis;
};

// The kd-tree the boxes come from, kept for point location (a new photon's first box, probe rays): inner node:
// split plane `v` of `axis`, children left (below) / right (above); leaf: axis = -1, left = box index.
struct RoomNode {
    float v;
    int32_t axis, left, right;
};
static_assert(sizeof(RoomNode) == 16, "RoomNode is one float4");

struct RoomScene {
    std::vector<RoomBox> boxes;
    std::vector<RoomBounds> bounds;            // per box
    std::vector<RoomFaceGrid> face_grids;
    std::vector<uint32_t> face_cells;          // the codes the grids index
    std::vector<RoomNode> nodes;               // nodes[0] = root
    std::vector<RoomStart> starts;             // per emitter (windows, then lights)
    float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
    int max_depth = 0;
    size_t kd_leaves = 0;                      // boxes before merging
    size_t face_parts = 0, wall_parts = 0;     // terminal codes over all faces; those that are colliders
    double build_ms = 0;
};

// Builds the decomposition.  Returns "" on success, else why the scene cannot use the room tier (an arbitrarily
// oriented collider, too many boxes): the caller then stays on the grid tier.
const char *build_rooms(RoomScene &out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows, int num_windows,
                        const fmgi_rect *lights, int num_lights);

// Host replay of the device traversal (tests/cpu, tools): wall index or -1, ray parameter, boxes crossed and face
// grids looked up.
int rooms_closest_hit(const RoomScene &rs, int box, const float o[3], const float d[3], float &t_out, int &box_out,
                      long &steps, long &tests);
// Box a ray that starts at p and travels along d is in (-1: outside the root box): tree descent, a point exactly
// on a split plane belongs to the side the ray travels towards.
int rooms_locate(const RoomScene &rs, const float p[3], const float d[3]);
// First box of a photon of `emitter` that starts at p (device: rooms_start): the emitter's partition, else the descent.
int rooms_start_box(const RoomScene &rs, int emitter, const float p[3], const float d[3]);

}  // namespace fmgi
