// Host-side builder of the room tier's box decomposition (room_tables.h).  Replaces, for this tier, both the
// reference's BSP build (photonmap.c:302-374) and the grid of scene_prep.cpp.
//
//   1. every collider becomes an axis-parallel rectangle record (normal axis, side, plane coordinate, extents);
//      a scene with an arbitrarily oriented collider is refused (the grid tier handles it);
//   2. kd-tree over the padded bounding box: a node whose open box still contains a piece of some collider is
//      split at a collider plane - big nodes at the plane nearest their middle, small ones at the plane that
//      covers the largest share of the node's cross-section - until nothing lies inside a leaf;
//   3. leaves are merged back: two boxes that share a whole face with no collider on it become one box (a split
//      plane runs through the whole node, also through the rooms its wall does not touch), to a fixed point;
//   4. per box and face: the colliders on that face whose normal points into the box and the boxes behind the face,
//      as a partition of the face stored as RoomFaceGrid records (a face with one thing on it is a code);
//   5. per emitter: the boxes its rectangle touches.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

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
    const auto zero3 = [](const float v[4]) { return v[0] == 0 && v[1] == 0 && v[2] == 0; };
    const auto nan3 = [](const float v[4]) { return v[0] != v[0] || v[1] != v[1] || v[2] != v[2]; };
    if (zero3(r.width) || zero3(r.height) || zero3(r.n) || nan3(r.width) || nan3(r.height) || nan3(r.n)) return 1;
    const int ai = single_axis3(r.width), aj = single_axis3(r.height), ak = single_axis3(r.n);
    if (ai < 0 || aj < 0 || ak < 0 || ai == aj || ak == ai || ak == aj) return 2;
    out.axis = ak; out.neg = r.n[ak] > 0 ? 0 : 1; out.id = id; out.c = r.pos[ak] + 0.0f;
    for (int k = 0; k < 3; k++) {
        const float a = r.pos[k], b = r.pos[k] + r.width[k] + r.height[k];
        out.lo[k] = fminf(a, b) + 0.0f; out.hi[k] = fmaxf(a, b) + 0.0f;          // + 0: no negative zeros
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

struct Node { int axis; float v; int left, right, leaf; };

struct Key3 {
    uint32_t k[3];
    bool operator==(const Key3 &o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; }
};
struct Key3Hash {
    size_t operator()(const Key3 &x) const
    {
        uint64_t h = x.k[0] * 0x9E3779B97F4A7C15ull;
        h = (h ^ (h >> 29)) + x.k[1] * 0xBF58476D1CE4E5B9ull;
        h = (h ^ (h >> 31)) + x.k[2] * 0x94D049BB133111EBull;
        return (size_t)(h ^ (h >> 32));
    }
};
Key3 key_of(const float p[3])
{
    Key3 k;
    for (int i = 0; i < 3; i++) { const float f = p[i] + 0.0f; memcpy(&k.k[i], &f, 4); }
    return k;
}

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

// ---- a face's partition as a 2-D kd-tree ----------------------------------------------------------------------

struct FaceItem {
    float lo[2], hi[2];          // in-plane extents (not clipped to the face)
    uint32_t code;               // kRoomCodeWall | wall index, or kRoomCodeBox | box index
};

struct Region { float lo[2], hi[2]; };

bool overlaps(const FaceItem &it, const Region &r)
{
    return fmaxf(it.lo[0], r.lo[0]) < fminf(it.hi[0], r.hi[0]) && fmaxf(it.lo[1], r.lo[1]) < fminf(it.hi[1], r.hi[1]);
}
bool covers(const FaceItem &it, const Region &r)
{
    return it.lo[0] <= r.lo[0] && it.hi[0] >= r.hi[0] && it.lo[1] <= r.lo[1] && it.hi[1] >= r.hi[1];
}

// A face (or an emitter rectangle) as RoomFaceGrid records: the edges of the things on it cut it into cells.
struct FaceGridBuilder {
    struct Out {
        std::vector<RoomFaceGrid> grids;
        std::vector<uint32_t> cells;
        size_t face_parts = 0, wall_parts = 0;
    };
    Out out;                     // grid codes / cell bases are local to `out`; the caller offsets them when it concatenates

    // The items of a region are pool[begin, end): colliders in wall-index order first, then portals; all overlap the
    // region with positive area.  The cells' lists are appended to the pool and dropped again when the cell is done -
    // no allocation per cell.
    std::vector<FaceItem> pool;
    std::vector<float> edges;
    uint32_t build(const Region &r, const std::vector<FaceItem> &items_in)
    {
        pool.assign(items_in.begin(), items_in.end());
        return build(r, 0, pool.size(), 0);
    }

    // up to three of the item edges strictly inside (lo, hi) along `ax`: all of them, or the quartiles
    int pick_splits(size_t begin, size_t end, int ax, float lo, float hi, float split[3])
    {
        edges.clear();
        for (size_t q = begin; q < end; q++) {
            const float a = pool[q].lo[ax], b = pool[q].hi[ax];
            if (a > lo && a < hi) edges.push_back(a);
            if (b > lo && b < hi) edges.push_back(b);
        }
        std::sort(edges.begin(), edges.end());
        edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
        const size_t n = edges.size();
        if (n <= 3) {
            for (size_t k = 0; k < n; k++) split[k] = edges[k];
            return (int)n;
        }
        split[0] = edges[n / 4]; split[1] = edges[n / 2]; split[2] = edges[(3 * n) / 4];
        return 3;
    }

    uint32_t build(const Region &r, size_t begin, size_t end, int depth)
    {
        if (begin == end) { out.face_parts++; return kRoomCodeMiss; }
        // the region belongs to the first item (lowest wall index; a collider hides the box behind it) if that covers it
        const FaceItem first = pool[begin];
        const bool first_is_wall = (first.code & kRoomCodeKind) == kRoomCodeWall;
        if (covers(first, r) && (first_is_wall || end - begin == 1)) {
            out.face_parts++;
            out.wall_parts += first_is_wall;
            return first.code;
        }
        float su[3] = {INFINITY, INFINITY, INFINITY}, sv[3] = {INFINITY, INFINITY, INFINITY};
        const int nu = pick_splits(begin, end, 0, r.lo[0], r.hi[0], su), nv = pick_splits(begin, end, 1, r.lo[1], r.hi[1], sv);
        if ((nu == 0 && nv == 0) || depth > 64) {     // nothing cuts the region, yet nothing covers it: keep the first
            out.face_parts++;
            out.wall_parts += first_is_wall;
            return first.code;
        }
        const uint32_t self = (uint32_t)out.grids.size(), base = (uint32_t)out.cells.size(), stride = (uint32_t)nu + 1;
        out.grids.push_back(RoomFaceGrid{{su[0], su[1], su[2]}, {sv[0], sv[1], sv[2]}, base, stride});
        out.cells.resize(out.cells.size() + (size_t)(nu + 1) * (nv + 1), kRoomCodeMiss);
        for (int iv = 0; iv <= nv; iv++)
            for (int iu = 0; iu <= nu; iu++) {
                Region c;
                c.lo[0] = iu == 0 ? r.lo[0] : su[iu - 1]; c.hi[0] = iu == nu ? r.hi[0] : su[iu];
                c.lo[1] = iv == 0 ? r.lo[1] : sv[iv - 1]; c.hi[1] = iv == nv ? r.hi[1] : sv[iv];
                const size_t mark = pool.size();
                for (size_t q = begin; q < end; q++)
                    if (overlaps(pool[q], c)) { const FaceItem it = pool[q]; pool.push_back(it); }
                const uint32_t code = build(c, mark, pool.size(), depth + 1);
                pool.resize(mark);
                out.cells[base + (uint32_t)iu + stride * (uint32_t)iv] = code;
            }
        return kRoomCodeNode | self;
    }
};

// ---- kd-tree over the colliders ------------------------------------------------------------------------------------

struct KdWork { int node; Box box; std::vector<int> ids; int depth; };
struct KdSubtree {
    int root_node = 0, max_depth = 0;
    std::vector<Node> nodes;                    // nodes[0] = the subtree's root; children are local indices
    std::vector<Box> boxes;
    std::vector<std::vector<int>> box_rects;
};

// The plane a node is split at; false: no collider lies inside the node's box (a leaf).
bool kd_choose_split(const std::vector<ARect> &rects, const KdWork &w, int &best_axis, float &best_c)
{
    best_axis = -1;
    best_c = 0;
    if (w.ids.size() > 96) {
        // big node: of the collider planes that cut it, the one nearest the middle of its longest side that has any
        int order[3] = {0, 1, 2};
        std::sort(order, order + 3, [&](int a, int b) { return w.box.hi[a] - w.box.lo[a] > w.box.hi[b] - w.box.lo[b]; });
        for (int oi = 0; oi < 3 && best_axis < 0; oi++) {
            const int a = order[oi];
            const float mid = 0.5f * (w.box.lo[a] + w.box.hi[a]);
            float best_d = INFINITY;
            for (int id : w.ids) {
                const ARect &r = rects[id];
                if (r.axis != a || !(fabsf(r.c - mid) < best_d) || !interior(w.box, r)) continue;
                best_d = fabsf(r.c - mid); best_axis = a; best_c = r.c;
            }
        }
        return best_axis >= 0;
    }
    // small node: score the collider planes that still cut it by the share of the cross-section they cover
    struct Plane { int axis; float c; double cover; };
    Plane planes[128];
    int np = 0;
    for (int id : w.ids) {
        const ARect &r = rects[id];
        if (!interior(w.box, r)) continue;
        double area = 1.0;
        for (int k = 0; k < 3; k++)
            if (k != r.axis) area *= (double)fminf(r.hi[k], w.box.hi[k]) - (double)fmaxf(r.lo[k], w.box.lo[k]);
        bool found = false;
        for (int q = 0; q < np; q++)
            if (planes[q].axis == r.axis && planes[q].c == r.c) { planes[q].cover += area; found = true; break; }
        if (!found) planes[np++] = Plane{r.axis, r.c, area};
    }
    double best_score = -1;
    for (int q = 0; q < np; q++) {
        const Plane &pl = planes[q];
        const int a = pl.axis;
        double cross = 1.0;
        for (int k = 0; k < 3; k++)
            if (k != a) cross *= (double)w.box.hi[k] - (double)w.box.lo[k];
        // coverage first; among equals the plane nearest the middle of the box
        const double mid = 1.0 - fabs(((double)pl.c - w.box.lo[a]) / ((double)w.box.hi[a] - w.box.lo[a]) - 0.5);
        const double score = pl.cover / cross + 1e-6 * mid;
        if (score > best_score || (score == best_score && (a < best_axis || (a == best_axis && pl.c < best_c)))) {
            best_score = score; best_axis = a; best_c = pl.c;
        }
    }
    return best_axis >= 0;
}

void kd_split(const std::vector<ARect> &rects, const KdWork &w, int axis, float c, KdWork &lw, KdWork &rw)
{
    lw.box = w.box; rw.box = w.box;
    lw.box.hi[axis] = c; rw.box.lo[axis] = c;
    lw.depth = rw.depth = w.depth + 1;
    lw.ids.reserve(w.ids.size()); rw.ids.reserve(w.ids.size());
    // every rectangle of the node touches the node's box: only the split axis decides which children it touches
    // (touches(): its plane within the closed range, or its extent overlapping the open one)
    const float blo = w.box.lo[axis], bhi = w.box.hi[axis];
    for (int id : w.ids) {
        const ARect &r = rects[id];
        bool left, right;
        if (r.axis == axis) { left = r.c <= c; right = r.c >= c; }
        else { left = fmaxf(r.lo[axis], blo) < fminf(r.hi[axis], c); right = fmaxf(r.lo[axis], c) < fminf(r.hi[axis], bhi); }
        if (left) lw.ids.push_back(id);
        if (right) rw.ids.push_back(id);
    }
}

// `leaves`: kd leaves of all subtrees so far; past `max_leaves` the build is given up (not a floor plan: rectangles
// floating in space cut each other into a number of boxes that grows much faster than their own number).
void kd_build_subtree(const std::vector<ARect> &rects, KdWork root, KdSubtree &out, std::atomic<size_t> &leaves, size_t max_leaves)
{
    out.root_node = root.node;
    root.node = 0;
    out.nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
    std::vector<KdWork> stack;
    stack.push_back(std::move(root));
    while (!stack.empty()) {
        KdWork w = std::move(stack.back());
        stack.pop_back();
        out.max_depth = std::max(out.max_depth, w.depth);
        int axis;
        float c;
        if (!kd_choose_split(rects, w, axis, c)) {
            if (leaves.fetch_add(1) >= max_leaves) return;
            out.nodes[w.node].axis = -1;
            out.nodes[w.node].leaf = (int)out.boxes.size();
            out.boxes.push_back(w.box);
            out.box_rects.push_back(std::move(w.ids));
            continue;
        }
        KdWork lw, rw;
        kd_split(rects, w, axis, c, lw, rw);
        lw.node = (int)out.nodes.size(); out.nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
        rw.node = (int)out.nodes.size(); out.nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
        out.nodes[w.node].axis = axis; out.nodes[w.node].v = c;
        out.nodes[w.node].left = lw.node; out.nodes[w.node].right = rw.node;
        stack.push_back(std::move(lw));
        stack.push_back(std::move(rw));
    }
}

// fn(i) for i in [0, n) on a pool of threads
template <typename Fn>
void run_parallel(size_t n, Fn fn, bool worth_threads = true)
{
    unsigned threads = worth_threads ? std::thread::hardware_concurrency() : 1u;
    if (const char *v = getenv("FMGI_BUILD_THREADS")) threads = (unsigned)atoi(v);
    threads = std::min<unsigned>(std::max(threads, 1u), 16u);
    threads = (unsigned)std::min<size_t>(threads, n);
    if (threads <= 1) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (std::thread &t : pool) t.join();
}

}  // namespace

const char *build_rooms(RoomScene &out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows, int num_windows,
                        const fmgi_rect *lights, int num_lights)
{
    const auto t0 = std::chrono::steady_clock::now();
    const bool timing = getenv("FMGI_ROOMS_TIMING") != nullptr;
    auto lap_t = t0;
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rooms]   %-28s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - lap_t).count());
        lap_t = now;
    };
    out = RoomScene();
    bool do_merge = true;
    if (const char *v = getenv("FMGI_ROOMS_MERGE")) do_merge = atoi(v) != 0;
    std::vector<ARect> rects;
    rects.reserve((size_t)num_walls);
    for (int i = 0; i < num_walls; i++) {
        ARect r;
        const int rc = to_axis_rect(walls[i], i, r);
        if (rc == 2) return "an arbitrarily oriented collider";
        if (rc == 0) rects.push_back(r);
    }
    if ((size_t)num_walls > kRoomCodeIndex) return "more than 2^30 walls";
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
    for (const ARect &r : rects)                 // (a degenerate wall cannot be hit and does not count)
        for (int k = 0; k < 3; k++) { root.lo[k] = fminf(root.lo[k], r.lo[k]); root.hi[k] = fmaxf(root.hi[k], r.hi[k]); }
    for (int i = 0; i < num_windows; i++) grow(windows[i]);
    for (int i = 0; i < num_lights; i++) grow(lights[i]);
    if (!(root.hi[0] >= root.lo[0])) for (int k = 0; k < 3; k++) { root.lo[k] = 0; root.hi[k] = 1; }
    for (int k = 0; k < 3; k++) { root.lo[k] -= 1.0f; root.hi[k] += 1.0f; }

    lap("rectangles + root box");
    // ---- kd-tree ------------------------------------------------------------------------------------------
    // The top of the tree is split here until there are a few dozen subtrees; those are independent and built by a
    // pool of threads into their own arrays, concatenated in subtree order (the result does not depend on the
    // number of threads).
    std::vector<Node> nodes;
    std::vector<Box> boxes;                     // kd leaves, later merged
    std::vector<std::vector<int>> box_rects;    // colliders that touch the box
    const size_t max_leaves = std::min<size_t>(8u << 20, 16 * rects.size() + 1024);
    std::atomic<size_t> leaves{0};
    {
        std::vector<KdWork> top;
        {
            KdWork w;
            w.node = 0; w.box = root; w.depth = 0;
            w.ids.resize(rects.size());
            for (size_t i = 0; i < rects.size(); i++) w.ids[i] = (int)i;
            nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
            top.push_back(std::move(w));
        }
        // level by level, the nodes of a level in parallel, until there are enough subtrees (or the nodes are small)
        const size_t want = rects.size() >= 4096 ? 64 : 1;
        std::vector<KdWork> small;
        struct Split { bool ok; int axis; float c; KdWork lw, rw; };
        while (!top.empty() && top.size() + small.size() < want) {
            std::vector<Split> sp(top.size());
            run_parallel(top.size(), [&](size_t i) {
                Split &x = sp[i];
                x.ok = top[i].ids.size() > 256 && kd_choose_split(rects, top[i], x.axis, x.c);
                if (x.ok) kd_split(rects, top[i], x.axis, x.c, x.lw, x.rw);
            });
            std::vector<KdWork> next;
            for (size_t i = 0; i < top.size(); i++) {
                Split &x = sp[i];
                if (!x.ok) { small.push_back(std::move(top[i])); continue; }
                x.lw.node = (int)nodes.size(); nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
                x.rw.node = (int)nodes.size(); nodes.push_back(Node{-1, 0.0f, -1, -1, -1});
                Node &n = nodes[top[i].node];
                n.axis = x.axis; n.v = x.c; n.left = x.lw.node; n.right = x.rw.node;
                next.push_back(std::move(x.lw));
                next.push_back(std::move(x.rw));
            }
            top = std::move(next);
        }
        for (KdWork &w : small) top.push_back(std::move(w));
        lap("kd top levels");
        std::vector<KdSubtree> subs(top.size());
        run_parallel(top.size(), [&](size_t i) { kd_build_subtree(rects, std::move(top[i]), subs[i], leaves, max_leaves); });
        if (leaves.load() > max_leaves) return "too many boxes for its number of colliders (not a floor plan)";
        lap("kd subtrees (parallel)");
        size_t more_nodes = 0, more_boxes = 0;
        for (const KdSubtree &st : subs) { more_nodes += st.nodes.size(); more_boxes += st.boxes.size(); }
        nodes.reserve(nodes.size() + more_nodes);
        boxes.reserve(more_boxes);
        box_rects.reserve(more_boxes);
        for (size_t i = 0; i < subs.size(); i++) {
            KdSubtree &st = subs[i];
            const int node_off = (int)nodes.size() - 1, leaf_off = (int)boxes.size();      // local node k > 0 -> node_off + k
            out.max_depth = std::max(out.max_depth, st.max_depth);
            auto fix = [&](Node n) {
                if (n.axis < 0) n.leaf += leaf_off;
                else { n.left += node_off; n.right += node_off; }
                return n;
            };
            nodes[st.root_node] = fix(st.nodes[0]);
            for (size_t k = 1; k < st.nodes.size(); k++) nodes.push_back(fix(st.nodes[k]));
            for (size_t k = 0; k < st.boxes.size(); k++) {
                boxes.push_back(st.boxes[k]);
                box_rects.push_back(std::move(st.box_rects[k]));
            }
        }
    }
    out.kd_leaves = boxes.size();
    lap("kd concatenate");
    const auto t_kd = std::chrono::steady_clock::now();

    // ---- merge boxes across collider-free shared faces ---------------------------------------------------------------
    const size_t nb0 = boxes.size();
    std::vector<int> merged_into(nb0);
    for (size_t i = 0; i < nb0; i++) merged_into[i] = (int)i;
    if (do_merge) {
        // box by lower corner: open addressing, never erased (a merged-away box is marked dead instead)
        size_t cap = 16;
        while (cap < nb0 * 2) cap *= 2;
        std::vector<Key3> slot_key(cap);
        std::vector<int> slot_box(cap, -1);
        auto find_box = [&](const Key3 &k) {
            for (size_t h = Key3Hash()(k) & (cap - 1);; h = (h + 1) & (cap - 1)) {
                if (slot_box[h] < 0) return -1;
                if (slot_key[h] == k) return slot_box[h];
            }
        };
        for (size_t i = 0; i < nb0; i++) {
            const Key3 k = key_of(boxes[i].lo);
            size_t h = Key3Hash()(k) & (cap - 1);
            while (slot_box[h] >= 0) h = (h + 1) & (cap - 1);
            slot_key[h] = k; slot_box[h] = (int)i;
        }
        std::vector<char> alive(nb0, 1);
        // Per box and axis: the box found at the far corner (-3: not looked up yet, -1: none - for good, since lower
        // corners never move and dead boxes stay dead) and the shape versions of the pair when it was last examined;
        // later rounds skip the pairs that have not changed.
        std::vector<int> nbr[3];
        std::vector<uint32_t> seen_i[3], seen_j[3], version(nb0, 1);
        for (int a = 0; a < 3; a++) { nbr[a].assign(nb0, -3); seen_i[a].assign(nb0, 0); seen_j[a].assign(nb0, 0); }
        bool changed = true;
        const int axis_order[3] = {2, 0, 1};
        while (changed) {
            changed = false;
            for (int ao = 0; ao < 3; ao++) {
                const int a = axis_order[ao], u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
                for (size_t i = 0; i < nb0; i++) {
                    if (!alive[i]) continue;
                    for (;;) {
                        Box &A = boxes[i];
                        int j = nbr[a][i];
                        if (j == -1) break;
                        if (j >= 0 && seen_i[a][i] == version[i] && seen_j[a][i] == version[j]) break;
                        if (j < 0) {
                            float corner[3] = {A.lo[0], A.lo[1], A.lo[2]};
                            corner[a] = A.hi[a];
                            j = find_box(key_of(corner));
                            if (j == (int)i) j = -1;
                            nbr[a][i] = j;
                        }
                        if (j < 0 || !alive[j]) { nbr[a][i] = -1; break; }
                        seen_i[a][i] = version[i]; seen_j[a][i] = version[j];
                        const Box &B = boxes[j];
                        if (B.hi[u] != A.hi[u] || B.hi[v] != A.hi[v]) break;
                        // a collider (facing either way) on the shared face keeps the boxes apart
                        bool wall = false;
                        for (int id : box_rects[i]) {
                            const ARect &r = rects[id];
                            if (r.axis == a && r.c == A.hi[a] && fmaxf(r.lo[u], A.lo[u]) < fminf(r.hi[u], A.hi[u]) &&
                                fmaxf(r.lo[v], A.lo[v]) < fminf(r.hi[v], A.hi[v])) { wall = true; break; }
                        }
                        if (wall) break;
                        A.hi[a] = B.hi[a];
                        box_rects[i].insert(box_rects[i].end(), box_rects[j].begin(), box_rects[j].end());
                        std::vector<int>().swap(box_rects[j]);
                        alive[j] = 0;
                        merged_into[j] = (int)i;
                        version[i]++;
                        nbr[a][i] = -3;                 // the far corner along this axis moved
                        changed = true;
                    }
                }
            }
        }
    }
    lap("merge");
    const auto t_merge = std::chrono::steady_clock::now();
    // final box ids
    std::vector<int> final_id(nb0, -1);
    int num_boxes = 0;
    for (size_t i = 0; i < nb0; i++)
        if (merged_into[i] == (int)i) final_id[i] = num_boxes++;
    auto resolve = [&](int leaf) {
        int r = leaf;
        while (merged_into[r] != r) r = merged_into[r];
        for (int q = leaf; merged_into[q] != q;) { const int nq = merged_into[q]; merged_into[q] = r; q = nq; }
        return r;
    };
    if ((size_t)num_boxes > kRoomCodeIndex) return "more than 2^30 boxes";
    out.boxes.resize((size_t)num_boxes);
    out.bounds.resize((size_t)num_boxes);

    // ---- faces (independent per box: chunks of boxes on the thread pool, face nodes concatenated in chunk order) ----------
    std::vector<int> root_of(nb0);
    for (size_t i = 0; i < nb0; i++) root_of[i] = resolve((int)i);
    std::vector<int> live;
    live.reserve((size_t)num_boxes);
    for (size_t i = 0; i < nb0; i++)
        if (final_id[i] >= 0) live.push_back((int)i);
    // the boxes around the building touch thousands of colliders and neighbours: they go first, so that no thread is
    // left alone with one at the end
    std::stable_sort(live.begin(), live.end(), [&](int a, int b) { return box_rects[a].size() > box_rects[b].size(); });
    lap("faces: order boxes");
    const size_t chunk = 4, num_chunks = (live.size() + chunk - 1) / chunk;
    std::vector<FaceGridBuilder::Out> chunk_out(num_chunks);
    // (a flat of a few hundred rectangles is built in half a millisecond: starting threads would cost more)
    run_parallel(num_chunks, [&](size_t ch) {
        FaceGridBuilder ftb;
        std::vector<int> behind;
        std::vector<FaceItem> items;
        for (size_t li = ch * chunk; li < std::min(live.size(), (ch + 1) * chunk); li++) {
            const int i = live[li];
            const Box &A = boxes[i];
            std::vector<int> &ids = box_rects[i];
            std::sort(ids.begin(), ids.end(), [&](int a, int b) { return rects[a].id < rects[b].id; });
            ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
            RoomBounds &bd = out.bounds[(size_t)final_id[i]];
            memset(&bd, 0, sizeof bd);
            bd.lo[0] = A.lo[0]; bd.lo[1] = A.lo[1]; bd.lo[2] = A.lo[2]; bd.hi_x = A.hi[0]; bd.hi_y = A.hi[1]; bd.hi_z = A.hi[2];
            uint32_t code[6];
            for (int f = 0; f < 6; f++) {
                const int a = f >> 1, side = f & 1;
                const float c = side ? A.hi[a] : A.lo[a];
                const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
                items.clear();
                for (int id : ids) {
                    const ARect &r = rects[id];
                    // leaving towards +axis (side 1) faces normals -axis, and the other way round (rectangle.c:70-72)
                    if (r.axis != a || r.c != c || r.neg != side) continue;
                    if (!(fmaxf(r.lo[u], A.lo[u]) < fminf(r.hi[u], A.hi[u])) || !(fmaxf(r.lo[v], A.lo[v]) < fminf(r.hi[v], A.hi[v])))
                        continue;
                    items.push_back(FaceItem{{r.lo[u], r.lo[v]}, {r.hi[u], r.hi[v]}, kRoomCodeWall | (uint32_t)r.id});
                }
                behind.clear();
                const bool at_root = side ? c >= root.hi[a] : c <= root.lo[a];
                if (!at_root) query_face(nodes, 0, a, c, side ? +1 : -1, A.lo, A.hi, behind);
                for (int &nb : behind) nb = root_of[nb];
                std::sort(behind.begin(), behind.end());
                behind.erase(std::unique(behind.begin(), behind.end()), behind.end());
                for (int nb : behind) {
                    const Box &B = boxes[nb];
                    items.push_back(FaceItem{{B.lo[u], B.lo[v]}, {B.hi[u], B.hi[v]}, kRoomCodeBox | (uint32_t)final_id[nb]});
                }
                const Region face = {{A.lo[u], A.lo[v]}, {A.hi[u], A.hi[v]}};
                code[f] = ftb.build(face, items);
            }
            RoomBox &rb = out.boxes[(size_t)final_id[i]];
            memset(&rb, 0, sizeof rb);
            for (int o = 0; o < 8; o++)
                for (int k = 0; k < 3; k++) {
                    const int side = (o >> k) & 1;
                    rb.oct[o].far[k] = side ? A.hi[k] : A.lo[k];
                    rb.oct[o].code[k] = code[2 * k + side];
                }
        }
        chunk_out[ch] = std::move(ftb.out);
    }, rects.size() >= 2048);
    lap("faces (parallel)");
    std::vector<size_t> grid_off(num_chunks + 1, 0), cell_off(num_chunks + 1, 0);
    for (size_t ch = 0; ch < num_chunks; ch++) {
        grid_off[ch + 1] = grid_off[ch] + chunk_out[ch].grids.size();
        cell_off[ch + 1] = cell_off[ch] + chunk_out[ch].cells.size();
        out.face_parts += chunk_out[ch].face_parts;
        out.wall_parts += chunk_out[ch].wall_parts;
    }
    out.face_grids.resize(grid_off[num_chunks]);
    out.face_cells.resize(cell_off[num_chunks]);
    const size_t groups = (num_chunks + 255) / 256;
    run_parallel(groups, [&](size_t gi) {
        for (size_t ch = gi * 256; ch < std::min(num_chunks, (gi + 1) * 256); ch++) {
            const uint32_t goff = (uint32_t)grid_off[ch], coff = (uint32_t)cell_off[ch];
            auto fix = [goff](uint32_t code) { return (code & kRoomCodeKind) == kRoomCodeNode ? code + goff : code; };
            RoomFaceGrid *gd = out.face_grids.data() + goff;
            for (RoomFaceGrid g : chunk_out[ch].grids) { g.base += coff; *gd++ = g; }
            uint32_t *cd = out.face_cells.data() + coff;
            for (uint32_t c : chunk_out[ch].cells) *cd++ = fix(c);
            for (size_t li = ch * chunk; li < std::min(live.size(), (ch + 1) * chunk); li++) {
                RoomBox &rb = out.boxes[(size_t)final_id[live[li]]];
                for (int o = 0; o < 8; o++)
                    for (int k = 0; k < 3; k++) rb.oct[o].code[k] = fix(rb.oct[o].code[k]);
            }
        }
    }, rects.size() >= 2048);
    lap("faces concatenate");
    const auto t_faces = std::chrono::steady_clock::now();
    // ---- per emitter: the boxes in front of its rectangle, as a partition of the rectangle (the same 2-D kd-trees as the
    // faces; an emitter in front of ONE box - a ceiling light, a window in its niche - is that box's code) -----------------
    {
        std::vector<int> stack_n;
        std::vector<int> seen;
        std::vector<FaceItem> items;
        FaceGridBuilder ftb;
        out.starts.assign((size_t)(num_windows + num_lights), RoomStart{kRoomCodeMiss, -1});
        for (int e = 0; e < num_windows + num_lights; e++) {
            ARect er;
            if (to_axis_rect(e < num_windows ? windows[e] : lights[e - num_windows], e, er) != 0) continue;   // tree descent
            const int a = er.axis, u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
            items.clear();
            seen.clear();
            stack_n.assign(1, 0);
            while (!stack_n.empty()) {
                const Node nd = nodes[stack_n.back()];
                stack_n.pop_back();
                if (nd.axis < 0) {
                    const int b = root_of[nd.leaf];
                    const Box &B = boxes[b];
                    // photons start at most 1e-5 off the plane, on the normal's side (er.neg: normal along -axis)
                    if (er.neg ? !(B.lo[a] < er.c && er.c <= B.hi[a]) : !(B.lo[a] <= er.c && er.c < B.hi[a])) continue;
                    if (!(fmaxf(B.lo[u], er.lo[u]) < fminf(B.hi[u], er.hi[u])) || !(fmaxf(B.lo[v], er.lo[v]) < fminf(B.hi[v], er.hi[v])))
                        continue;
                    if (std::find(seen.begin(), seen.end(), b) != seen.end()) continue;
                    seen.push_back(b);
                    items.push_back(FaceItem{{B.lo[u], B.lo[v]}, {B.hi[u], B.hi[v]}, kRoomCodeBox | (uint32_t)final_id[b]});
                    continue;
                }
                if (nd.axis == a) {
                    if (er.c < nd.v || (er.c == nd.v && er.neg)) stack_n.push_back(nd.left);
                    if (er.c > nd.v || (er.c == nd.v && !er.neg)) stack_n.push_back(nd.right);
                } else {
                    if (er.lo[nd.axis] < nd.v) stack_n.push_back(nd.left);
                    if (er.hi[nd.axis] > nd.v) stack_n.push_back(nd.right);
                }
            }
            const Region region = {{er.lo[u], er.lo[v]}, {er.hi[u], er.hi[v]}};
            ftb.out = FaceGridBuilder::Out();
            const uint32_t code = ftb.build(region, items);
            // the emitters' records go behind the faces' records
            const uint32_t goff = (uint32_t)out.face_grids.size(), coff = (uint32_t)out.face_cells.size();
            auto fix = [goff](uint32_t c) { return (c & kRoomCodeKind) == kRoomCodeNode ? c + goff : c; };
            for (RoomFaceGrid g : ftb.out.grids) { g.base += coff; out.face_grids.push_back(g); }
            for (uint32_t c : ftb.out.cells) out.face_cells.push_back(fix(c));
            out.starts[(size_t)e] = RoomStart{fix(code), a};
        }
        if (out.starts.empty()) out.starts.push_back(RoomStart{kRoomCodeMiss, -1});
    }
    if (out.face_grids.size() > kRoomCodeIndex) return "more than 2^30 face records";
    if (out.face_grids.empty()) out.face_grids.push_back(RoomFaceGrid{{INFINITY, INFINITY, INFINITY}, {INFINITY, INFINITY, INFINITY}, 0u, 1u});
    if (out.face_cells.empty()) out.face_cells.push_back(kRoomCodeMiss);          // the device tables are never empty
    // the tree itself, for point location
    out.nodes.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) {
        RoomNode &n = out.nodes[i];
        n.v = nodes[i].v; n.axis = nodes[i].axis;
        n.left = nodes[i].axis < 0 ? final_id[resolve(nodes[i].leaf)] : nodes[i].left;
        n.right = nodes[i].right;
    }
    for (int k = 0; k < 3; k++) { out.root_lo[k] = root.lo[k]; out.root_hi[k] = root.hi[k]; }
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("FMGI_ROOMS_TIMING")) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[rooms] kd %.2f ms, merge %.2f ms, faces %.2f ms, emitters + tree %.2f ms\n", ms(t0, t_kd), ms(t_kd, t_merge),
                ms(t_merge, t_faces), ms(t_faces, std::chrono::steady_clock::now()));
    }
    return "";
}

// Box a ray that starts at p and travels along d is in: tree descent; a point exactly on a split plane belongs to
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

// the cell of grid record `g` the point (pu, pv) falls into (device: the same six compares)
static uint32_t rooms_grid_lookup(const RoomScene &rs, uint32_t g, float pu, float pv)
{
    const RoomFaceGrid &G = rs.face_grids[g];
    const uint32_t iu = (pu >= G.su[0]) + (pu >= G.su[1]) + (pu >= G.su[2]);
    const uint32_t iv = (pv >= G.sv[0]) + (pv >= G.sv[1]) + (pv >= G.sv[2]);
    return rs.face_cells[G.base + iu + G.stride * iv];
}

int rooms_start_box(const RoomScene &rs, int emitter, const float p[3], const float d[3])
{
    const RoomStart &st = rs.starts[(size_t)emitter];
    uint32_t code = st.code;
    const float pu = st.axis == 0 ? p[1] : p[0], pv = st.axis == 2 ? p[1] : p[2];
    while ((code & kRoomCodeKind) == kRoomCodeNode) code = rooms_grid_lookup(rs, code, pu, pv);
    if ((code & kRoomCodeKind) == kRoomCodeBox) return (int)(code & kRoomCodeIndex);
    return rooms_locate(rs, p, d);
}

// Host replay of the device traversal (rooms_walk in trace_kernels.cuh), same float operations.
int rooms_closest_hit(const RoomScene &rs, int box, const float o[3], const float d[3], float &t_out, int &box_out,
                      long &steps, long &tests)
{
    t_out = INFINITY;
    box_out = box;
    const int oct = (d[0] > 0 ? 1 : 0) | (d[1] > 0 ? 2 : 0) | (d[2] > 0 ? 4 : 0);
    float inv[3];
    for (int k = 0; k < 3; k++) inv[k] = d[k] == 0 ? -1e30f : 1.0f / d[k];
    for (int guard = 0; guard < 1 << 16 && box >= 0; guard++) {
        const RoomOctant &R = rs.boxes[(size_t)box].oct[oct];
        steps++;
        float tk[3];
        for (int k = 0; k < 3; k++) tk[k] = (R.far[k] - o[k]) * inv[k];
        int a = 0;
        if (tk[1] < tk[0]) a = 1;
        if (tk[2] < fminf(tk[0], tk[1])) a = 2;
        const float t = fminf(fminf(tk[0], tk[1]), tk[2]);
        const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
        const float pu = fmaf(t, d[u], o[u]), pv = fmaf(t, d[v], o[v]);
        uint32_t code = R.code[a];
        while ((code & kRoomCodeKind) == kRoomCodeNode) {
            tests++;
            code = rooms_grid_lookup(rs, code, pu, pv);
        }
        const uint32_t kind = code & kRoomCodeKind, index = code & kRoomCodeIndex;
        if (kind == kRoomCodeWall) {
            if (!(t >= 0)) {                          // the plane lies behind the origin: not a hit (rectangle.c:76) -
                box = rooms_locate(rs, o, d);         // the origin is in another box: go on from there
                continue;
            }
            t_out = (R.far[a] - o[a]) / d[a];
            box_out = box;
            return (int)index;
        }
        if (kind == kRoomCodeMiss) return -1;
        box = (int)index;
    }
    return -1;
}

}  // namespace fmgi
