// Host-side scene precompute: Rectangle soup -> traversal and shading tables.
//
// Compiled with -ffp-contract=off: the per-rectangle constants that feed the texel index
// (width/|width|, |width| ...) must be the very floats the reference recomputes on every call of
// getTileIdAt (rectangle.c:205-230) and the sampler basis must be the one
// getCosineDistributedRandomRay builds (vector3_cl.c:139-144), so each expression below keeps
// the reference's operation order: length() = sqrtf(x*x + y*y + z*z) (vector3_cl.c:93),
// div_vec3() = multiply by 1.0f/len (vector3_cl.c:53-58), normalized() likewise (:95-100).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "scene_tables.h"

namespace fmgi {

namespace {

struct V3 { float x, y, z; };
inline V3 ld(const float *p) { return {p[0], p[1], p[2]}; }
inline float length(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 scale(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 normalized(V3 a) { float fac = 1.0f / length(a); return scale(a, fac); }
inline void st(float *d, V3 a) { d[0] = a.x; d[1] = a.y; d[2] = a.z; }

// vector3_cl.c:139-144 (identical in getDiffuseSkyRandomRay :118-123)
void sampler_basis(V3 n, V3 &u, V3 &v)
{
    u = {0, 0, 1};
    if (fabs(dot(u, n)) >= 0.999999f)
        u = {0, 1, 0};
    v = normalized(cross(u, n));
    u = normalized(cross(v, n));
}

int single_axis(V3 a)
{
    int nz = (a.x != 0) + (a.y != 0) + (a.z != 0);
    if (nz != 1) return -1;
    return a.x != 0 ? 0 : (a.y != 0 ? 1 : 2);
}

inline float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

bool finite3(V3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }

}  // namespace

void sampler_basis(const float n[3], float u[3], float v[3])
{
    V3 uu, vv;
    sampler_basis(ld(n), uu, vv);
    st(u, uu); st(v, vv);
}

uint64_t photon_budget(float area, int samples_per_area)
{
    float n = samples_per_area * area;      // int * float -> float, photonmap.c:418
    if (!(n > 0)) return 0;
    return (uint64_t)n;
}

const char *prepare_scene(HostScene &out, const fmgi_rect *walls, int num_walls,
                          const fmgi_rect *windows, int num_windows,
                          const fmgi_rect *lights, int num_lights, int num_texels)
{
    out = HostScene();
    out.num_walls = num_walls; out.num_windows = num_windows; out.num_lights = num_lights;
    out.num_texels = num_texels;
    if (num_walls < 0 || num_windows < 0 || num_lights < 0 || num_texels < 0)
        return "negative count";

    struct Axis1 { float c, mid_i, half_i, mid_j, half_j; int id; };
    std::vector<Axis1> groups[6];         // 2*k + (normal sign > 0 ? 0 : 1)
    out.shade.resize(num_walls);

    for (int r = 0; r < num_walls; r++) {
        const fmgi_rect &w = walls[r];
        V3 pos = ld(w.pos), wd = ld(w.width), ht = ld(w.height), n = ld(w.n);
        if (!finite3(pos) || !finite3(wd) || !finite3(ht) || !finite3(n))
            return "non-finite wall rectangle";
        int base = w.lightmap[0], tw = w.lightmap[1], th = w.lightmap[2];
        if (tw < 1 || th < 1 || base < 0 || (int64_t)base + (int64_t)tw * th > (int64_t)num_texels)
            return "wall lightmap tile range lies outside the atlas";

        float wlen = length(wd), hlen = length(ht);
        ShadeRect &s = out.shade[r];
        memset(&s, 0, sizeof s);
        V3 u, v;
        sampler_basis(n, u, v);
        if (tw > 32767 || th > 32767)
            return "more than 32767 lightmap tiles along one edge of a wall";
        st(s.pos, pos); s.base = base;
        st(s.wn, scale(wd, 1.0f / wlen)); s.wlen = wlen;
        st(s.hn, scale(ht, 1.0f / hlen)); s.hlen = hlen;
        st(s.n, n); s.tiles = tw | (th << 16);
        st(s.u, u);
        st(s.v, v);

        if (!(wlen > 0) || !(hlen > 0) || !(length(n) > 0))
            continue;                       // degenerate: zero area, can never be hit

        int ai = single_axis(wd), aj = single_axis(ht), ak = single_axis(n);
        if (ai >= 0 && aj >= 0 && ak >= 0 && ai != aj && ak != ai && ak != aj) {
            // in-plane axes in ascending order, whichever of width/height they belong to
            int i = ai < aj ? ai : aj, j = ai < aj ? aj : ai;
            V3 far = {pos.x + wd.x + ht.x, pos.y + wd.y + ht.y, pos.z + wd.z + ht.z};
            const float lo_i = fminf(comp(pos, i), comp(far, i)), hi_i = fmaxf(comp(pos, i), comp(far, i));
            const float lo_j = fminf(comp(pos, j), comp(far, j)), hi_j = fmaxf(comp(pos, j), comp(far, j));
            Axis1 a;
            a.c = comp(pos, ak);
            a.mid_i = 0.5f * (lo_i + hi_i); a.half_i = 0.5f * (hi_i - lo_i);
            a.mid_j = 0.5f * (lo_j + hi_j); a.half_j = 0.5f * (hi_j - lo_j);
            a.id = r;
            groups[2 * ak + (comp(n, ak) > 0 ? 0 : 1)].push_back(a);
            out.num_axis_rects++;
        } else {
            GeneralRect g;
            g.nx = n.x; g.ny = n.y; g.nz = n.z; g.nd = dot(n, pos);
            V3 wn = scale(wd, 1.0f / wlen), hn = scale(ht, 1.0f / hlen);
            g.wx = wn.x; g.wy = wn.y; g.wz = wn.z; g.wlen = wlen;
            g.hx = hn.x; g.hy = hn.y; g.hz = hn.z; g.hlen = hlen;
            g.px = pos.x; g.py = pos.y; g.pz = pos.z; g.id = r;
            out.general.push_back(g);
        }
    }
    // interleave the P and M list of every axis in blocks of two rectangles, padded with
    // entries that can never be hit (plane coordinate NaN)
    const float nanv = std::nanf("");
    const Axis1 pad = {nanv, 0.0f, -1.0f, 0.0f, -1.0f, -1};
    for (int k = 0; k < 3; k++) {
        const std::vector<Axis1> &P = groups[2 * k], &M = groups[2 * k + 1];
        const size_t len = std::max(P.size(), M.size());
        const size_t pairs = (len + 1) / 2;
        out.pair_begin[k] = (int)(out.axis.size() / 2);
        for (size_t j = 0; j < pairs; j++)
            for (int sgn = 0; sgn < 2; sgn++) {
                const std::vector<Axis1> &L = sgn ? M : P;
                const Axis1 &a = 2 * j < L.size() ? L[2 * j] : pad;
                const Axis1 &b = 2 * j + 1 < L.size() ? L[2 * j + 1] : pad;
                AxisPairBlock blk = {a.c, a.mid_i, a.half_i, a.mid_j, b.c, b.mid_i, b.half_i, b.mid_j,
                                     a.half_j, b.half_j, a.id, b.id};
                out.axis.push_back(blk);
            }
    }
    out.pair_begin[3] = (int)(out.axis.size() / 2);

    // emitters: windows first, then lights (photonmap.c:412-431)
    for (int e = 0; e < num_windows + num_lights; e++) {
        const fmgi_rect &w = e < num_windows ? windows[e] : lights[e - num_windows];
        V3 n = ld(w.n);
        if (!finite3(ld(w.pos)) || !finite3(ld(w.width)) || !finite3(ld(w.height)) || !finite3(n))
            return "non-finite emitter rectangle";
        EmitterRec er;
        memset(&er, 0, sizeof er);
        V3 u, v;
        sampler_basis(n, u, v);
        st(er.pos, ld(w.pos)); er.is_window = e < num_windows;
        st(er.width, ld(w.width)); st(er.height, ld(w.height));
        st(er.n, n); st(er.u, u); st(er.v, v);
        out.emitters.push_back(er);
        out.emitter_area.push_back(length(ld(w.width)) * length(ld(w.height)));   // photonmap.c:417
    }
    return "";
}

// ---- grid tier --------------------------------------------------------------------------------------

void build_grid(HostScene &out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows, int num_windows,
                const fmgi_rect *lights, int num_lights, float cell_hint)
{
    std::vector<GridItem> items;
    grid_classify(out, walls, num_walls, windows, num_windows, lights, num_lights, cell_hint, items);
    grid_assemble_host(out, items);
}

// First half of build_grid: the grid's geometry and plane tables (GridDesc) and, per collider, which list(s) it goes
// to, the cells it covers and its record.  O(number of rectangles); the per-cell work is grid_assemble_host / the
// device assembler (grid_build.cuh).
void grid_classify(HostScene &out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows, int num_windows,
                   const fmgi_rect *lights, int num_lights, float cell_hint, std::vector<GridItem> &items)
{
    GridDesc &g = out.grid;
    g = GridDesc();
    out.grid_ranges.clear();
    out.grid_recs.clear();
    out.grid_table.clear();
    out.grid_overflow_horizontal = 0;
    out.grid_misc = 0;

    // bounding box of everything a ray can start from or hit
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    auto grow = [&](const fmgi_rect &r) {
        for (int c = 0; c < 4; c++) {
            const float x = r.pos[0] + (c & 1 ? r.width[0] : 0.0f) + (c & 2 ? r.height[0] : 0.0f);
            const float y = r.pos[1] + (c & 1 ? r.width[1] : 0.0f) + (c & 2 ? r.height[1] : 0.0f);
            xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
        }
    };
    for (int i = 0; i < num_walls; i++) grow(walls[i]);
    for (int i = 0; i < num_windows; i++) grow(windows[i]);
    for (int i = 0; i < num_lights; i++) grow(lights[i]);
    if (!(xmax >= xmin)) { xmin = ymin = 0; xmax = ymax = 1; }

    const float w = fmaxf(xmax - xmin, 1e-3f), h = fmaxf(ymax - ymin, 1e-3f);
    float cell = cell_hint;
    if (!(cell > 0)) cell = sqrtf(w * h / fmaxf(1.0f, 0.5f * (float)num_walls));
    // keep the grid below 4M cells whatever the hint
    while ((double)(w / cell + 5) * (double)(h / cell + 5) > 4.0e6) cell *= 1.5f;
    g.cell = cell; g.inv_cell = 1.0f / cell;
    // two cells of margin all around: the walk stops where the ray leaves the box that excludes the
    // outermost ring (GridWalk::t_exit), so a step taken on a rounding error still lands on a cell
    g.x0 = xmin - 2 * cell; g.y0 = ymin - 2 * cell;
    g.nx = (int)ceilf(w / cell) + 4; g.ny = (int)ceilf(h / cell) + 4;
    const int ncell = g.nx * g.ny;

    // classify: which list does each wall go to, and with which record
    typedef GridItem Item;
    items.clear();
    items.reserve((size_t)num_walls);
    std::vector<float> up, down;
    const float eps = 1e-3f * cell;
    int general_index = 0;
    float zlo = INFINITY, zhi = -INFINITY;
    for (int r = 0; r < num_walls; r++) {
        const fmgi_rect &q = walls[r];
        V3 pos = ld(q.pos), wd = ld(q.width), ht = ld(q.height), n = ld(q.n);
        if (!(length(wd) > 0) || !(length(ht) > 0) || !(length(n) > 0)) continue;     // as prepare_scene
        Item it;
        memset(&it, 0, sizeof it);
        const int ai = single_axis(wd), aj = single_axis(ht), ak = single_axis(n);
        const bool axis = ai >= 0 && aj >= 0 && ak >= 0 && ai != aj && ak != ai && ak != aj;
        V3 far = {pos.x + wd.x + ht.x, pos.y + wd.y + ht.y, pos.z + wd.z + ht.z};
        it.list = -1;                                       // -1: walk list
        if (axis) {
            const int i = ai < aj ? ai : aj, j = ai < aj ? aj : ai;
            const float lo_i = fminf(comp(pos, i), comp(far, i)), hi_i = fmaxf(comp(pos, i), comp(far, i));
            const float lo_j = fminf(comp(pos, j), comp(far, j)), hi_j = fmaxf(comp(pos, j), comp(far, j));
            it.rec.c = comp(pos, ak);
            it.rec.mid_i = 0.5f * (lo_i + hi_i); it.rec.half_i = 0.5f * (hi_i - lo_i);
            it.rec.mid_j = 0.5f * (lo_j + hi_j); it.rec.half_j = 0.5f * (hi_j - lo_j);
            const int neg = comp(n, ak) > 0 ? 0 : 1;
            it.axis = ak; it.neg = neg;
            it.rec.tag = (uint32_t)r | (ak == 1 ? kTagAlongY : 0u);
            if (ak == 2) {                                  // horizontal: try the plane table
                std::vector<float> &pl = neg ? down : up;
                size_t p = 0;
                while (p < pl.size() && pl[p] != it.rec.c) p++;
                if (p == pl.size() && pl.size() < (size_t)kMaxPlanesPerSign) pl.push_back(it.rec.c);
                if (p < pl.size()) {
                    it.list = (neg ? kMaxPlanesPerSign : 0) + (int)p;
                    it.rec.tag = (uint32_t)r | kTagHorizontal;
                } else {
                    out.grid_overflow_horizontal++;
                    it.rec.tag = (uint32_t)r | kTagMisc | (neg ? kTagNegative : 0u);
                }
            }
        } else {
            it.axis = 3; it.neg = 0;
            it.rec.tag = (uint32_t)general_index | kTagMisc | kTagHorizontal;
        }
        if (!axis) general_index++;                         // same order as HostScene::general
        if (it.rec.tag & kTagMisc) out.grid_misc++;
        // (x, y) bounding box over the four corners
        float bx0 = INFINITY, bx1 = -INFINITY, by0 = INFINITY, by1 = -INFINITY;
        for (int c = 0; c < 4; c++) {
            const float x = pos.x + (c & 1 ? wd.x : 0.0f) + (c & 2 ? ht.x : 0.0f);
            const float y = pos.y + (c & 1 ? wd.y : 0.0f) + (c & 2 ? ht.y : 0.0f);
            bx0 = fminf(bx0, x); bx1 = fmaxf(bx1, x); by0 = fminf(by0, y); by1 = fmaxf(by1, y);
        }
        auto cellx = [&](float x) { int c = (int)floorf((x - g.x0) * g.inv_cell); return c < 0 ? 0 : (c >= g.nx ? g.nx - 1 : c); };
        auto celly = [&](float y) { int c = (int)floorf((y - g.y0) * g.inv_cell); return c < 0 ? 0 : (c >= g.ny ? g.ny - 1 : c); };
        if (it.list < 0)
            for (int c = 0; c < 4; c++) {
                const float z = pos.z + (c & 1 ? wd.z : 0.0f) + (c & 2 ? ht.z : 0.0f);
                zlo = fminf(zlo, z); zhi = fmaxf(zhi, z);
            }
        it.cx0 = cellx(bx0 - eps); it.cx1 = cellx(bx1 + eps);
        it.cy0 = celly(by0 - eps); it.cy1 = celly(by1 + eps);
        items.push_back(it);
    }
    // Plane order: a ray travelling down meets the upward-facing planes from the top, a ray travelling up
    // the downward-facing ones from the bottom; with the nearest plane first a hit bounds the later ones
    // away (t < best fails) before their cell is looked up.
    {
        std::vector<int> remap(2 * kMaxPlanesPerSign, -1);
        for (int sgn = 0; sgn < 2; sgn++) {
            std::vector<float> &pl = sgn ? down : up;
            std::vector<int> order(pl.size());
            for (size_t p = 0; p < pl.size(); p++) order[p] = (int)p;
            std::sort(order.begin(), order.end(), [&](int a, int b) { return sgn ? pl[a] < pl[b] : pl[a] > pl[b]; });
            std::vector<float> sorted(pl.size());
            for (size_t p = 0; p < pl.size(); p++) {
                sorted[p] = pl[order[p]];
                remap[sgn * kMaxPlanesPerSign + order[p]] = sgn * kMaxPlanesPerSign + (int)p;
            }
            pl = sorted;
        }
        for (Item &it : items)
            if (it.list >= 0) it.list = remap[it.list];
    }
    g.planes_up = (int)up.size(); g.planes_down = (int)down.size();
    for (size_t p = 0; p < up.size(); p++) g.plane_z[p] = up[p];
    for (size_t p = 0; p < down.size(); p++) g.plane_z[kMaxPlanesPerSign + p] = down[p];
    g.ncell = ncell;
    if (!(zhi >= zlo)) { zlo = 0.0f; zhi = 0.0f; }
    g.wall_z_lo = zlo; g.wall_z_hi = zhi;
    g.planes_max = std::max(g.planes_up, g.planes_down);
    g.bx = -g.x0 * g.inv_cell; g.by = -g.y0 * g.inv_cell;
    g.exit_lo_x = g.x0 + g.cell; g.exit_hi_x = g.x0 + (float)(g.nx - 1) * g.cell;
    g.exit_lo_y = g.y0 + g.cell; g.exit_hi_y = g.y0 + (float)(g.ny - 1) * g.cell;
    // list bases of T and the fast plane tables
    g.down_base = g.planes_up * ncell;
    g.walk_base = (g.planes_up + g.planes_down) * ncell;
    for (int dir = 0; dir < 2; dir++)
        for (int i = 0; i < 4; i++) {
            const int count = dir == 0 ? g.planes_up : g.planes_down;
            g.fast_z[dir][i] = i < count ? g.plane_z[(dir == 0 ? 0 : kMaxPlanesPerSign) + i] : std::nanf("");
            g.fast_base[dir][i] = (dir == 0 ? 0 : g.down_base) + (i < count ? i : 0) * ncell;
        }
}

// Second half of build_grid on the host: bins the items into per-(list, cell) lists and lays out T.
void grid_assemble_host(HostScene &out, const std::vector<GridItem> &items)
{
    typedef GridItem Item;
    GridDesc &g = out.grid;
    const int ncell = g.ncell;
    const float cell = g.cell;
    out.grid_ranges.clear();
    out.grid_recs.clear();

    // lists are numbered: [0, 8) planes +z, [8, 16) planes -z, 16 + combo = walk lists, where
    // combo = (d.x > 0) + 2 * (d.y > 0) is the sign combination of the rays that walk the list
    const int num_lists = kNumGridLists;
    auto for_each_list = [&](const Item &it, auto &&fn) {
        if (it.list >= 0) { fn(it.list); return; }
        const int k = it.axis, neg = it.neg;
        for (int combo = 0; combo < 4; combo++) {
            // normal +x is faced by d.x < 0 (combo bit 0 clear), normal -x by d.x > 0; same for y
            if (k == 0 && ((combo & 1) != 0) != (neg != 0)) continue;
            if (k == 1 && ((combo & 2) != 0) != (neg != 0)) continue;
            fn(kWalkListBase + combo);
        }
    };
    std::vector<int32_t> count((size_t)num_lists * ncell + 1, 0);
    for (const Item &it : items)
        for_each_list(it, [&](int l) {
            for (int cy = it.cy0; cy <= it.cy1; cy++)
                for (int cx = it.cx0; cx <= it.cx1; cx++)
                    count[(size_t)l * ncell + cy * g.nx + cx]++;
        });
    std::vector<int32_t> begin((size_t)num_lists * ncell + 1, 0);
    for (size_t i = 0; i < (size_t)num_lists * ncell; i++) begin[i + 1] = begin[i] + count[i];
    out.grid_recs.resize((size_t)begin[(size_t)num_lists * ncell]);
    std::vector<int32_t> fill(begin.begin(), begin.end() - 1);
    for (const Item &it : items)
        for_each_list(it, [&](int l) {
            for (int cy = it.cy0; cy <= it.cy1; cy++)
                for (int cx = it.cx0; cx <= it.cx1; cx++)
                    out.grid_recs[(size_t)fill[(size_t)l * ncell + cy * g.nx + cx]++] = it.rec;
        });
    // plane lists: the lookup stops at the first rectangle that contains the crossing point, so put
    // the rectangle that covers most of the cell first (floors tile the plane: usually one dominates)
    for (int l = 0; l < kWalkListBase; l++)
        for (int cy = 0; cy < g.ny; cy++)
            for (int cx = 0; cx < g.nx; cx++) {
                const size_t idx = (size_t)l * ncell + cy * g.nx + cx;
                const int b = begin[idx], e = begin[idx + 1];
                if (e - b < 2) continue;
                const float x0c = g.x0 + cx * cell, y0c = g.y0 + cy * cell;
                auto overlap = [&](const GridRec &r) {
                    const float ox = fminf(r.mid_i + r.half_i, x0c + cell) - fmaxf(r.mid_i - r.half_i, x0c);
                    const float oy = fminf(r.mid_j + r.half_j, y0c + cell) - fmaxf(r.mid_j - r.half_j, y0c);
                    return fmaxf(ox, 0.0f) * fmaxf(oy, 0.0f);
                };
                std::stable_sort(out.grid_recs.begin() + b, out.grid_recs.begin() + e,
                                 [&](const GridRec &a, const GridRec &c) { return overlap(a) > overlap(c); });
            }
    out.grid_ranges.resize((size_t)2 * num_lists * ncell);
    for (size_t i = 0; i < (size_t)num_lists * ncell; i++) {
        out.grid_ranges[2 * i] = begin[i];
        out.grid_ranges[2 * i + 1] = begin[i + 1];
    }

    // T: heads (first record inline + continuation range), then the remaining records
    const int num_used = g.planes_up + g.planes_down + 4;
    auto used_list = [&](int u) {      // compact list number -> CSR list number
        if (u < g.planes_up) return u;
        if (u < g.planes_up + g.planes_down) return kMaxPlanesPerSign + (u - g.planes_up);
        return kWalkListBase + (u - g.planes_up - g.planes_down);
    };
    GridRec dummy;
    memset(&dummy, 0, sizeof dummy);
    dummy.half_i = -1.0f; dummy.half_j = -1.0f; dummy.c = std::nanf("");
    out.grid_table.assign((size_t)num_used * ncell, dummy);
    for (int u = 0; u < num_used; u++) {
        const int l = used_list(u);
        for (int cell = 0; cell < ncell; cell++) {
            const int b = begin[(size_t)l * ncell + cell], e = begin[(size_t)l * ncell + cell + 1];
            if (b == e) continue;
            // every record says what follows it: [next, end) of its list, empty after the last one
            GridRec head = out.grid_recs[b];
            head.next = (int32_t)out.grid_table.size();
            head.end = head.next + (e - b - 1);
            for (int q = b + 1; q < e; q++) {
                GridRec rest = out.grid_recs[q];
                rest.next = (int32_t)out.grid_table.size() + 1;
                rest.end = head.end;
                out.grid_table.push_back(rest);
            }
            out.grid_table[(size_t)u * ncell + cell] = head;
        }
    }
}

}  // namespace fmgi
