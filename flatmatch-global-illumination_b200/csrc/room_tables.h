// Room tier: closest hit by walking a box decomposition of the flat (built by rooms_build.cpp).
//
// Every collider parseLayout.c emits is an axis-parallel rectangle, and a flat is mostly empty boxes: rooms,
// door and window niches, the inside of walls.  A kd-tree splits the scene's bounding box at collider planes
// until NO collider lies inside a leaf box: all of them sit on leaf faces.  A ray then never tests a collider
// "on spec": it leaves its box through one of the three faces it travels towards, and the face's entry list says
// what is there - a collider that faces into the box (hit: the closest one by construction), or a neighbouring
// box (portal: continue there), or nothing (the ray leaves the scene).  A ray in a room does ONE step (its own
// room's walls), a ray through a doorway three; no grid walk through empty cells, no per-cell candidate tests, no
// separate pass over the horizontal planes (floors, ceilings, sills and lintels are faces like any other).
// Back-face culling (rectangle.c:70-72) is structural: a face lists only the colliders whose normal points into
// its box.  Ties (rectangle edges shared by two colliders) go to the lowest wall index, as the reference's strict
// `<` does (photonmap.cl:199): the colliders of a face come first and in index order.
#pragma once
#include <stdint.h>
#include <vector>

#include "../../include/fmgi.h"

namespace fmgi {

// One leaf box.  Face f = 2 * axis + (1 if the ray leaves towards +axis): its entries are
// E[face_begin[f] .. face_begin[f + 1]).
struct RoomLeaf {
    float lo[3], hi[3];
    int32_t pad0[2];
    int32_t face_begin[7];
    int32_t pad1;
};
static_assert(sizeof(RoomLeaf) == 64, "RoomLeaf is four float4");

// One rectangle on a leaf face, in the face's two in-plane axes (ascending axis order).
struct RoomEntry {
    float u_lo, u_hi, v_lo, v_hi;
    int32_t target;          // >= 0: wall index (a hit); < 0: ~(index of the leaf behind the face) (a portal)
    float c;                 // the face's plane coordinate
    int32_t pad[2];
};
static_assert(sizeof(RoomEntry) == 32, "RoomEntry is two float4");

// The kd-tree the leaves come from, kept for point location (a new photon's first leaf, probe rays): inner node:
// split plane `v` of `axis`, children left (below) / right (above); leaf: axis = -1, left = leaf index.
struct RoomNode {
    float v;
    int32_t axis, left, right;
};
static_assert(sizeof(RoomNode) == 16, "RoomNode is one float4");

struct RoomScene {
    std::vector<RoomLeaf> leaves;
    std::vector<RoomEntry> entries;
    std::vector<RoomNode> nodes;               // nodes[0] = root
    // where an emitter's photons start: the leaves that touch the emitter rectangle (windows, then lights):
    // start_leaves[start_range[2e] .. start_range[2e + 1]) - one or two boxes, checked by containment; the tree
    // descent is the fallback
    std::vector<int32_t> start_range;
    std::vector<int32_t> start_leaves;
    float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
    int max_depth = 0;
    double build_ms = 0;
};

// Builds the decomposition.  Returns "" on success, else why the scene cannot use the room tier (an arbitrarily
// oriented collider, too many leaves): the caller then stays on the grid tier.
const char *build_rooms(RoomScene &out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows, int num_windows,
                        const fmgi_rect *lights, int num_lights);

// Host replay of the device traversal (tests/cpu, tools): wall index or -1, ray parameter, steps and entries tested.
int rooms_closest_hit(const RoomScene &rs, int leaf, const float o[3], const float d[3], float &t_out, int &leaf_out,
                      long &steps, long &tests);
// Leaf a ray that starts at p and travels along d is in (-1: outside the root box): tree descent, a point exactly
// on a split plane belongs to the side the ray travels towards.
int rooms_locate(const RoomScene &rs, const float p[3], const float d[3]);

}  // namespace fmgi
