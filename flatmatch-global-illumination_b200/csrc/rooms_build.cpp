// Host-side builder of the room tier's box decomposition (room_tables.h).  Replaces, for this tier, both the
// reference's BSP build (photonmap.c:302-374) and the grid of scene_prep.cpp.
//
//   1. every collider becomes an axis-parallel rectangle record (normal axis, side, plane coordinate, extents);
//      a scene with an arbitrarily oriented collider is refused (the grid tier handles it);
//   2. kd-tree over the padded bounding box: a node whose open box still contains a piece of some collider is
//      split at the collider plane that covers the largest share of the node's cross-section (room-separating
//      walls, floor and ceiling first; sills and lintels last, when the node is already the niche they sit in);
//   3. per leaf and face: the colliders on that face whose normal points into the box, in wall-index order, then the
//      leaves behind the face (found by a tree query of the face rectangle);
//   4. per emitter: nothing - photons locate their first leaf by a tree descent (RoomNode) on the device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>

#include "room_tables.h"

namespace fmgi {

namespace {

struct ARect {
    int axis, neg, id;
    float c;
    float lo[3], hi[3];
};

struct Box { float lo[3], hi[3]; };

int single_axis3(const float v[4])
{
    const int nz = (v[0] != 0) + (v[1] != 0) + (v[2] != 0);
    if (nz != 1) return -1;
    return v[0] != 0 ? 0 : (v[1] != 0 ? 1 : 2);
}

// 0: ok, 1: degenerate (zero area: can never be hit), 2: not axis parallel
int to_axis_rect(const fmgi_rect &r, int id, ARect &out)
{
    const float wl = sqrtf(r.width[0] * r.width[0] + r.width[1] * r.width[1] + r.width[2] * r.width[2]);
    const float hl = sqrtf(r.height[0] * r.height[0] + r.height[1] * r.height[1] + r.height[2] * r.height[2]);
    const float nl = sqrtf(r.n[0] * r.n[0] + r.n[1] * r.n[1] + r.n[2] * r.n[2]);
    if (!(wl > 0) || !(hl > 0) || !(nl > 0)) return 1;
    const int ai = single_axis3(r.width), aj = single_axis3(r.height), ak = single_axis3(r.n);
    if (ai < 0 || aj < 0 || ak < 0 || ai == aj || ak == ai || ak == aj) return 2;
    out.axis = ak; out.neg = r.n[ak] > 0 ? 0 : 1; out.id = id; out.c = r.pos[ak];
    for (int k = 0; k < 3; k++) {
        const float a = r.pos[k], b = r.pos[k] + r.width[k] + r.height[k];
        out.lo[k] = fminf(a, b); out.hi[k] = fmaxf(a, b);
    }
    out.lo[ak] = out.hi[ak] = out.c;
    return 0;
}

// a piece of positive area of r lies strictly inside the box
bool interior(const Box &b, const ARect &r)
{
    const int a = r.axis;
    if (!(r.c > b.lo[a] && r.c < b.hi[a])) return false;
    for (int k = 0; k < 3; k++)
        if (k != a && !(fmaxf(r.lo[k], b.lo[k]) < fminf(r.hi[k], b.hi[k]))) return false;
    return true;
}

// r has a piece of positive area inside or on the boundary of the box
bool touches(const Box &b, const ARect &r)
{
    const int a = r.axis;
    if (!(r.c >= b.lo[a] && r.c <= b.hi[a])) return false;
    for (int k = 0; k < 3; k++)
        if (k != a && !(fmaxf(r.lo[k], b.lo[k]) < fminf(r.hi[k], b.hi[k]))) return false;
    return true;
}

struct Node { int axis; float v; int left, right, leaf; };

}  // namespace

namespace {

// leaves whose boxes lie just behind plane coordinate c (on side `beyond`: +1 higher, -1 lower) of axis a and overlap
// the open rectangle [qlo, qhi] of the other two axes with positive area
void query_face(const std::vector<Node> &nodes, int n, int a, float c, int beyond, const float qlo[3], const float qhi[3],
                std::vector<int> &out)
{
    const Node &nd = nodes[n];
    if (nd.axis < 0) { out.push_back(nd.leaf); return; }
    const int s = nd.axis;
    if (s == a) {
        if (c == nd.v) query_face(nodes, beyond > 0 ? nd.right : nd.left, a, c, beyond, qlo, qhi, out);
        else if (c < nd.v) query_face(nodes, nd.left, a, c, beyond, qlo, qhi, out);
        else query_face(nodes, nd.right, a, c, beyond, qlo, qhi, out);
    } else {
        if (qlo[s] < nd.v) query_face(nodes, nd.left, a, c, beyond, qlo, qhi, out);
        if (qhi[s] > nd.v) query_face(nodes, nd.right, a, c, beyond, qlo, qhi, out);
    }
}

}  // namespace

const char *build_rooms(RoomScene &out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows, int num_windows,
                        const fmgi_rect *lights, int num_lights)
{
    const auto t0 = std::chrono::steady_clock::now();
    out = RoomScene();
    std::vector<ARect> rects;
    rects.reserve((size_t)num_walls);
    for (int i = 0; i < num_walls; i++) {
        ARect r;
        const int rc = to_axis_rect(walls[i], i, r);
        if (rc == 2) return "an arbitrarily oriented collider";
        if (rc == 0) rects.push_back(r);
    }
    // padded bounding box of everything a ray can start from or hit
    Box root;
    for (int k = 0; k < 3; k++) { root.lo[k] = INFINITY; root.hi[k] = -INFINITY; }
    auto grow = [&](const fmgi_rect &r) {
        for (int c = 0; c < 4; c++)
            for (int k = 0; k < 3; k++) {
                const float x = r.pos[k] + (c & 1 ? r.width[k] : 0.0f) + (c & 2 ? r.height[k] : 0.0f);
                root.lo[k] = fminf(root.lo[k], x); root.hi[k] = fmaxf(root.hi[k], x);
            }
    };
    for (int i = 0; i < num_walls; i++) grow(walls[i]);
    for (int i = 0; i < num_windows; i++) grow(windows[i]);
    for (int i = 0; i < num_lights; i++) grow(lights[i]);
    if (!(root.hi[0] >= root.lo[0])) for (int k = 0; k < 3; k++) { root.lo[k] = 0; root.hi[k] = 1; }
    for (int k = 0; k < 3; k++) { root.lo[k] -= 1.0f; root.hi[k] += 1.0f; }

    // ---- kd-tree ------------------------------------------------------------------------------------------
    std::vector<Node> nodes;
    struct Work { int node; Box box; std::vector<int> ids; int depth; };
    std::vector<Work> stack;
    std::vector<std::vector<int>> leaf_rects;
    {
        Work w;
        w.node = 0; w.box = root; w.depth = 0;
        w.ids.resize(rects.size());
        for (size_t i = 0; i < rects.size(); i++) w.ids[i] = (int)i;
        nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
        stack.push_back(std::move(w));
    }
    const size_t max_leaves = 8u << 20;
    while (!stack.empty()) {
        Work w = std::move(stack.back());
        stack.pop_back();
        out.max_depth = std::max(out.max_depth, w.depth);
        // score the collider planes that still cut this box
        std::map<std::pair<int, float>, double> cover;
        for (int id : w.ids) {
            const ARect &r = rects[id];
            if (!interior(w.box, r)) continue;
            double area = 1.0;
            for (int k = 0; k < 3; k++)
                if (k != r.axis) area *= (double)fminf(r.hi[k], w.box.hi[k]) - (double)fmaxf(r.lo[k], w.box.lo[k]);
            cover[{r.axis, r.c}] += area;
        }
        if (cover.empty()) {
            const int leaf = (int)out.leaves.size();
            if ((size_t)leaf >= max_leaves) return "more than 8M leaf boxes";
            RoomLeaf L;
            memset(&L, 0, sizeof L);
            for (int k = 0; k < 3; k++) { L.lo[k] = w.box.lo[k]; L.hi[k] = w.box.hi[k]; }
            out.leaves.push_back(L);
            leaf_rects.push_back(std::move(w.ids));
            nodes[w.node].axis = -1;
            nodes[w.node].leaf = leaf;
            continue;
        }
        int best_axis = -1;
        float best_c = 0;
        double best_score = -1;
        for (const auto &kv : cover) {
            const int a = kv.first.first;
            double cross = 1.0;
            for (int k = 0; k < 3; k++)
                if (k != a) cross *= (double)w.box.hi[k] - (double)w.box.lo[k];
            // coverage first; among equals the plane nearest the middle of the box
            const double mid = 1.0 - fabs(((double)kv.first.second - w.box.lo[a]) / ((double)w.box.hi[a] - w.box.lo[a]) - 0.5);
            const double score = kv.second / cross + 1e-6 * mid;
            if (score > best_score) { best_score = score; best_axis = a; best_c = kv.first.second; }
        }
        Work lw, rw;
        lw.box = w.box; rw.box = w.box;
        lw.box.hi[best_axis] = best_c; rw.box.lo[best_axis] = best_c;
        lw.depth = rw.depth = w.depth + 1;
        for (int id : w.ids) {
            if (touches(lw.box, rects[id])) lw.ids.push_back(id);
            if (touches(rw.box, rects[id])) rw.ids.push_back(id);
        }
        lw.node = (int)nodes.size(); nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
        rw.node = (int)nodes.size(); nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
        nodes[w.node].axis = best_axis; nodes[w.node].v = best_c;
        nodes[w.node].left = lw.node; nodes[w.node].right = rw.node;
        stack.push_back(std::move(lw));
        stack.push_back(std::move(rw));
    }

    // ---- face lists: colliders facing into the box (index order), then the leaves behind the face ------------------
    std::vector<int> behind;
    for (size_t li = 0; li < out.leaves.size(); li++) {
        RoomLeaf &L = out.leaves[li];
        std::sort(leaf_rects[li].begin(), leaf_rects[li].end(), [&](int a, int b) { return rects[a].id < rects[b].id; });
        for (int f = 0; f < 6; f++) {
            const int a = f >> 1, side = f & 1;
            const float c = side ? L.hi[a] : L.lo[a];
            const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
            L.face_begin[f] = (int32_t)out.entries.size();
            for (int id : leaf_rects[li]) {
                const ARect &r = rects[id];
                // leaving towards +axis (side 1) faces normals -axis, and the other way round (rectangle.c:70-72)
                if (r.axis != a || r.c != c || r.neg != side) continue;
                if (!(fmaxf(r.lo[u], L.lo[u]) < fminf(r.hi[u], L.hi[u])) || !(fmaxf(r.lo[v], L.lo[v]) < fminf(r.hi[v], L.hi[v])))
                    continue;
                RoomEntry e;
                memset(&e, 0, sizeof e);
                e.u_lo = r.lo[u]; e.u_hi = r.hi[u]; e.v_lo = r.lo[v]; e.v_hi = r.hi[v];
                e.target = r.id; e.c = c;
                out.entries.push_back(e);
            }
            behind.clear();
            const bool at_root = side ? c >= root.hi[a] : c <= root.lo[a];
            if (!at_root) query_face(nodes, 0, a, c, side ? +1 : -1, L.lo, L.hi, behind);
            for (int nb : behind) {
                const RoomLeaf &B = out.leaves[nb];
                RoomEntry e;
                memset(&e, 0, sizeof e);
                e.u_lo = B.lo[u]; e.u_hi = B.hi[u]; e.v_lo = B.lo[v]; e.v_hi = B.hi[v];
                e.target = ~nb; e.c = c;
                out.entries.push_back(e);
            }
        }
        L.face_begin[6] = (int32_t)out.entries.size();
    }
    // per emitter: the leaves its rectangle touches (closed overlap with the rectangle grown by the start offset)
    {
        std::vector<int> stack_n;
        for (int e = 0; e < num_windows + num_lights; e++) {
            const fmgi_rect &r = e < num_windows ? windows[e] : lights[e - num_windows];
            float lo[3], hi[3];
            for (int k = 0; k < 3; k++) {
                const float a = r.pos[k], b2 = r.pos[k] + r.width[k] + r.height[k];
                lo[k] = fminf(a, b2) - 3e-5f; hi[k] = fmaxf(a, b2) + 3e-5f;
            }
            out.start_range.push_back((int32_t)out.start_leaves.size());
            stack_n.assign(1, 0);
            while (!stack_n.empty()) {
                const Node nd = nodes[stack_n.back()];
                stack_n.pop_back();
                if (nd.axis < 0) { out.start_leaves.push_back(nd.leaf); continue; }
                if (lo[nd.axis] <= nd.v) stack_n.push_back(nd.left);
                if (hi[nd.axis] >= nd.v) stack_n.push_back(nd.right);
            }
            out.start_range.push_back((int32_t)out.start_leaves.size());
        }
    }
    // the tree itself, for point location
    out.nodes.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) {
        RoomNode &n = out.nodes[i];
        n.v = nodes[i].v; n.axis = nodes[i].axis;
        n.left = nodes[i].axis < 0 ? nodes[i].leaf : nodes[i].left;
        n.right = nodes[i].right;
    }
    for (int k = 0; k < 3; k++) { out.root_lo[k] = root.lo[k]; out.root_hi[k] = root.hi[k]; }
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return "";
}

// Leaf a ray that starts at p and travels along d is in: tree descent; a point exactly on a split plane belongs to
// the side the ray travels towards.
int rooms_locate(const RoomScene &rs, const float p[3], const float d[3])
{
    for (int k = 0; k < 3; k++)
        if (!(p[k] >= rs.root_lo[k] && p[k] <= rs.root_hi[k])) return -1;
    int n = 0;
    while (rs.nodes[n].axis >= 0) {
        const RoomNode &nd = rs.nodes[n];
        const float x = p[nd.axis];
        const bool right = x > nd.v || (x == nd.v && d[nd.axis] > 0);
        n = right ? nd.right : nd.left;
    }
    return rs.nodes[n].left;
}

// Host replay of the device traversal.
int rooms_closest_hit(const RoomScene &rs, int leaf, const float o[3], const float d[3], float &t_out, int &leaf_out,
                      long &steps, long &tests)
{
    const float inf = INFINITY;
    t_out = inf;
    leaf_out = leaf;
    for (int guard = 0; guard < 1 << 16 && leaf >= 0; guard++) {
        const RoomLeaf &L = rs.leaves[leaf];
        steps++;
        float tk[3];
        for (int k = 0; k < 3; k++) tk[k] = d[k] == 0 ? inf : ((d[k] > 0 ? L.hi[k] : L.lo[k]) - o[k]) / d[k];
        int a = 0;
        if (tk[1] < tk[a]) a = 1;
        if (tk[2] < tk[a]) a = 2;
        const float t = tk[a];
        if (!(t < inf)) return -1;
        const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
        const float pu = o[u] + t * d[u], pv = o[v] + t * d[v];
        const int f = 2 * a + (d[a] > 0 ? 1 : 0);
        int next = -1;
        bool found = false;
        for (int q = L.face_begin[f]; q < L.face_begin[f + 1]; q++) {
            const RoomEntry &e = rs.entries[q];
            tests++;
            if (pu >= e.u_lo && pu <= e.u_hi && pv >= e.v_lo && pv <= e.v_hi) {
                if (e.target >= 0) { t_out = t; leaf_out = leaf; return e.target; }
                next = ~e.target;
                found = true;
                break;
            }
        }
        if (!found) return -1;
        leaf = next;
    }
    return -1;
}

}  // namespace fmgi
