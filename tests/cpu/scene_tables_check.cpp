// TEST INFRASTRUCTURE (CPU): checks the host-built scene tables of the grid tier (csrc/scene_prep.cpp)
// without a GPU.  Built and run by tests/test_scene_tables_cpu.py.
//
//   1. structure of the head table T against the CSR lists it is derived from (heads, dummies, continuation
//      ranges, plane order);
//   2. a host replay of the device walk over T (the semantics of GridWalk in csrc/trace_kernels.cuh: walls by the
//      2-D DDA first, then the z planes crossed before the wall hit) against a brute-force scan with the
//      reference's intersects() semantics (rectangle.c:67-95: back-face culling, t >= 0, edges inclusive,
//      strict "<" so the lowest index wins ties) on random rays.
//
// usage: scene_tables_check scene.bin num_rays [cell]      (scene.bin as written by the test: int32 counts,
//        then the walls / windows / lights tables of 80-byte Rectangle records)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../flatmatch-global-illumination_b200/csrc/scene_tables.h"

using namespace fmgi;

static uint64_t rng_state = 0x1234567ull;
static inline double urand()
{
    rng_state += 0x9E3779B97F4A7C15ull;
    uint64_t z = rng_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

static int fails = 0;
#define CHECK(cond, ...)                                                      \
    do {                                                                      \
        if (!(cond)) {                                                        \
            if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } \
            fails++;                                                          \
        }                                                                     \
    } while (0)

// rectangle.c:67-95 in float, the reference's operation order
static float intersects(const fmgi_rect &r, const float o[3], const float d[3], float closest)
{
    const float denom = r.n[0] * d[0] + r.n[1] * d[1] + r.n[2] * d[2];
    if (denom >= 0) return -1;
    const float fac = (r.n[0] * (r.pos[0] - o[0]) + r.n[1] * (r.pos[1] - o[1]) + r.n[2] * (r.pos[2] - o[2])) / denom;
    if (fac < 0) return -1;
    if (!(fac < closest)) return -1;
    const float p[3] = {o[0] + d[0] * fac - r.pos[0], o[1] + d[1] * fac - r.pos[1], o[2] + d[2] * fac - r.pos[2]};
    const float wl = sqrtf(r.width[0] * r.width[0] + r.width[1] * r.width[1] + r.width[2] * r.width[2]);
    const float hl = sqrtf(r.height[0] * r.height[0] + r.height[1] * r.height[1] + r.height[2] * r.height[2]);
    if (!(wl > 0) || !(hl > 0)) return -1;
    const float iw = 1.0f / wl, ih = 1.0f / hl;
    const float u = (r.width[0] * iw) * p[0] + (r.width[1] * iw) * p[1] + (r.width[2] * iw) * p[2];
    const float v = (r.height[0] * ih) * p[0] + (r.height[1] * ih) * p[1] + (r.height[2] * ih) * p[2];
    if (u < 0 || v < 0 || u > wl || v > hl) return -1;
    return fac;
}

struct Replay {
    const HostScene &sc;
    const std::vector<fmgi_rect> &walls;
    explicit Replay(const HostScene &s, const std::vector<fmgi_rect> &w) : sc(s), walls(w) {}

    bool test_misc(const GridRec &rec, const float o[3], const float d[3], float best, float &t_out) const
    {
        if (rec.tag & kTagHorizontal) {       // arbitrarily oriented: the full test on the caller's rectangle
            const GeneralRect &g = sc.general[rec.tag & kTagIdMask];
            const float t = intersects(walls[g.id], o, d, best);
            if (t < 0) return false;
            t_out = t;
            return true;
        }
        const bool facing = (rec.tag & kTagNegative) ? d[2] > 0 : d[2] < 0;
        const float t = (rec.c - o[2]) / d[2];
        const float pi = t * d[0] + o[0] - rec.mid_i, pj = t * d[1] + o[1] - rec.mid_j;
        if (!(facing && t >= 0 && t < best && fabsf(pi) <= rec.half_i && fabsf(pj) <= rec.half_j)) return false;
        t_out = t;
        return true;
    }

    // wall index (-1: miss) and distance, walking T the way the device does
    int closest(const float o[3], const float d[3], float &t_out, long &steps) const
    {
        const GridDesc &g = sc.grid;
        const std::vector<GridRec> &T = sc.grid_table;
        const float inf = INFINITY;
        float best = inf;
        int win = -1;
        // ---- walls
        {
            const bool xp = d[0] > 0, yp = d[1] > 0, x0 = d[0] == 0, y0 = d[1] == 0;
            const float ix = 1.0f / d[0], iy = 1.0f / d[1], iz = 1.0f / d[2];
            int cx = (int)floorf(o[0] * g.inv_cell + g.bx), cy = (int)floorf(o[1] * g.inv_cell + g.by);
            cx = std::min(std::max(cx, 0), g.nx - 1); cy = std::min(std::max(cy, 0), g.ny - 1);
            float tmx = x0 ? inf : (((float)(cx + (xp ? 1 : 0)) * g.cell + g.x0) - o[0]) * ix;
            float tmy = y0 ? inf : (((float)(cy + (yp ? 1 : 0)) * g.cell + g.y0) - o[1]) * iy;
            const float ex = x0 ? inf : ((xp ? g.exit_hi_x : g.exit_lo_x) - o[0]) * ix;
            const float ey = y0 ? inf : ((yp ? g.exit_hi_y : g.exit_lo_y) - o[1]) * iy;
            const float ez = d[2] == 0 ? inf : ((d[2] < 0 ? g.wall_z_lo : g.wall_z_hi) - o[2]) * iz * 1.0001f;
            const float t_exit = fminf(fminf(ex, ey), ez);
            const int sx = xp ? 1 : -1, sy = yp ? g.nx : -g.nx;
            int ci = g.walk_base + ((xp ? 1 : 0) + (yp ? 2 : 0)) * g.ncell + cy * g.nx + cx;
            int cur = ci;
            best = fminf(best, t_exit);
            for (;;) {
                steps++;
                const GridRec &rec = T[cur];
                if (rec.c == rec.c) {               // not a dummy head
                    if (rec.tag & kTagMisc) {
                        float t;
                        if (test_misc(rec, o, d, best, t)) { best = t; win = cur; }
                    } else {
                        const bool ky = (rec.tag & kTagAlongY) != 0;
                        if (!(ky ? y0 : x0)) {
                            const float t = (rec.c - (ky ? o[1] : o[0])) * (ky ? iy : ix);
                            const float pi = t * (ky ? d[0] : d[1]) + (ky ? o[0] : o[1]) - rec.mid_i;
                            const float pj = t * d[2] + o[2] - rec.mid_j;
                            if (t >= 0 && t < best && fabsf(pi) <= rec.half_i && fabsf(pj) <= rec.half_j) { best = t; win = cur; }
                        }
                    }
                }
                if (rec.next < rec.end) { cur = rec.next; continue; }
                const float t_next = fminf(tmx, tmy);
                if (!(t_next < best)) break;
                if (tmx < tmy) { ci += sx; tmx += g.cell * fabsf(ix); } else { ci += sy; tmy += g.cell * fabsf(iy); }
                if (ci < g.walk_base || ci >= g.walk_base + 4 * g.ncell) { CHECK(false, "walk left the table"); break; }
                cur = ci;
            }
            if (win < 0) best = inf;
        }
        // ---- planes crossed before the wall hit, nearest first
        {
            const bool down = d[2] < 0;
            const int count = down ? g.planes_up : g.planes_down;
            const int base = down ? 0 : g.down_base;
            for (int pl = 0; pl < count && d[2] != 0; pl++) {
                const float z = g.plane_z[(down ? 0 : kMaxPlanesPerSign) + pl];
                const float t = (z - o[2]) / d[2];
                if (!(t >= 0 && t < best)) continue;
                const float x = t * d[0] + o[0], y = t * d[1] + o[1];
                const int px = (int)floorf(x * g.inv_cell + g.bx), py = (int)floorf(y * g.inv_cell + g.by);
                if (px < 0 || py < 0 || px >= g.nx || py >= g.ny) continue;
                int q = base + pl * g.ncell + py * g.nx + px;
                for (;;) {
                    steps++;
                    const GridRec &rec = T[q];
                    if (fabsf(x - rec.mid_i) <= rec.half_i && fabsf(y - rec.mid_j) <= rec.half_j) { best = t; win = q; break; }
                    if (q < g.walk_base + 4 * g.ncell) {        // a head: continue with its list
                        if (rec.next >= rec.end) break;
                        q = rec.next;
                    } else {
                        if (rec.next >= rec.end) break;
                        q = rec.next;
                    }
                }
            }
        }
        t_out = best;
        if (win < 0) return -1;
        const GridRec &w = T[win];
        if ((w.tag & (kTagMisc | kTagHorizontal)) == (kTagMisc | kTagHorizontal)) return sc.general[w.tag & kTagIdMask].id;
        return (int)(w.tag & kTagIdMask);
    }
};

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.bin num_rays [cell]\n", argv[0]); return 2; }
    const int num_rays = atoi(argv[2]);
    const float cell = argc > 3 ? (float)atof(argv[3]) : 0.0f;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("scene"); return 2; }
    int32_t cnt[3];
    if (fread(cnt, 4, 3, f) != 3) return 2;
    std::vector<fmgi_rect> walls(cnt[0]), windows(cnt[1]), lights(cnt[2]);
    if (fread(walls.data(), sizeof(fmgi_rect), cnt[0], f) != (size_t)cnt[0]) return 2;
    if (fread(windows.data(), sizeof(fmgi_rect), cnt[1], f) != (size_t)cnt[1]) return 2;
    if (fread(lights.data(), sizeof(fmgi_rect), cnt[2], f) != (size_t)cnt[2]) return 2;
    fclose(f);

    HostScene sc;
    const char *why = prepare_scene(sc, walls.data(), cnt[0], windows.data(), cnt[1], lights.data(), cnt[2], 1 << 30);
    if (why[0]) { printf("prepare_scene: %s\n", why); return 1; }
    build_grid(sc, walls.data(), cnt[0], windows.data(), cnt[1], lights.data(), cnt[2], cell);
    const GridDesc &g = sc.grid;
    const std::vector<GridRec> &T = sc.grid_table;
    const int num_used = g.planes_up + g.planes_down + 4;
    printf("grid %dx%d cell %.3f planes %d/%d T %zu records (%d heads) misc %d\n", g.nx, g.ny, g.cell, g.planes_up,
           g.planes_down, T.size(), num_used * g.ncell, sc.grid_misc);

    // ---- 1. structure ---------------------------------------------------------------------------------
    CHECK(g.ncell == g.nx * g.ny, "ncell");
    CHECK(g.down_base == g.planes_up * g.ncell && g.walk_base == (g.planes_up + g.planes_down) * g.ncell, "list bases");
    CHECK(T.size() >= (size_t)num_used * g.ncell, "table shorter than its heads");
    for (int p = 1; p < g.planes_up; p++) CHECK(g.plane_z[p] < g.plane_z[p - 1], "up planes not highest-first");
    for (int p = 1; p < g.planes_down; p++)
        CHECK(g.plane_z[kMaxPlanesPerSign + p] > g.plane_z[kMaxPlanesPerSign + p - 1], "down planes not lowest-first");
    const int32_t *ranges = sc.grid_ranges.data();
    size_t rest_seen = 0;
    for (int u = 0; u < num_used; u++) {
        const int l = u < g.planes_up ? u : (u < g.planes_up + g.planes_down ? kMaxPlanesPerSign + (u - g.planes_up)
                                                                             : kWalkListBase + (u - g.planes_up - g.planes_down));
        for (int cell_i = 0; cell_i < g.ncell; cell_i++) {
            const int b = ranges[2 * ((size_t)l * g.ncell + cell_i)], e = ranges[2 * ((size_t)l * g.ncell + cell_i) + 1];
            const GridRec &head = T[(size_t)u * g.ncell + cell_i];
            if (b == e) {
                CHECK(head.c != head.c && head.half_i < 0 && head.half_j < 0 && head.next >= head.end, "empty list without a dummy head");
                continue;
            }
            CHECK(memcmp(&head, &sc.grid_recs[b], 24) == 0, "head differs from the list's first record");
            CHECK(head.end - head.next == e - b - 1, "head continuation length");
            CHECK(head.next >= num_used * g.ncell && (size_t)head.end <= T.size(), "continuation outside the record area");
            for (int q = b + 1, t = head.next; q < e; q++, t++) {
                CHECK(memcmp(&T[t], &sc.grid_recs[q], 24) == 0, "rest record differs");
                CHECK(T[t].next == t + 1 && T[t].end == head.end, "rest record continuation");
                rest_seen++;
            }
        }
    }
    CHECK(rest_seen + (size_t)num_used * g.ncell == T.size(), "table size: %zu rest records + heads != %zu", rest_seen, T.size());
    // every collider sits in some list of every cell its centre falls into
    for (int r = 0; r < cnt[0]; r++) {
        const fmgi_rect &w = walls[r];
        const float cxw = w.pos[0] + 0.5f * (w.width[0] + w.height[0]), cyw = w.pos[1] + 0.5f * (w.width[1] + w.height[1]);
        const float wl = sqrtf(w.width[0] * w.width[0] + w.width[1] * w.width[1] + w.width[2] * w.width[2]);
        const float hl = sqrtf(w.height[0] * w.height[0] + w.height[1] * w.height[1] + w.height[2] * w.height[2]);
        if (!(wl > 0) || !(hl > 0)) continue;
        const int px = (int)floorf(cxw * g.inv_cell + g.bx), py = (int)floorf(cyw * g.inv_cell + g.by);
        // two cells of margin all around (one less where a centre sits on the box edge and rounds down)
        CHECK(px >= 1 && py >= 1 && px < g.nx - 2 && py < g.ny - 2, "wall centre not inside the grid's margin");
        bool found = false;
        for (int u = 0; u < num_used && !found; u++) {
            int q = u * g.ncell + py * g.nx + px;
            for (;;) {
                const GridRec &rec = T[q];
                const bool general = (rec.tag & (kTagMisc | kTagHorizontal)) == (kTagMisc | kTagHorizontal);
                if (rec.c == rec.c && (general ? sc.general[rec.tag & kTagIdMask].id : (int)(rec.tag & kTagIdMask)) == r) { found = true; break; }
                if (rec.next >= rec.end) break;
                q = rec.next;
            }
        }
        CHECK(found, "wall %d is in no list of the cell of its centre", r);
    }

    // ---- 2. host replay of the walk against brute force ------------------------------------------------
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
    for (const fmgi_rect &w : walls)
        for (int c = 0; c < 4; c++) {
            const float p[3] = {w.pos[0] + (c & 1 ? w.width[0] : 0) + (c & 2 ? w.height[0] : 0),
                                w.pos[1] + (c & 1 ? w.width[1] : 0) + (c & 2 ? w.height[1] : 0),
                                w.pos[2] + (c & 1 ? w.width[2] : 0) + (c & 2 ? w.height[2] : 0)};
            xmin = fminf(xmin, p[0]); xmax = fmaxf(xmax, p[0]); ymin = fminf(ymin, p[1]); ymax = fmaxf(ymax, p[1]);
            zmin = fminf(zmin, p[2]); zmax = fmaxf(zmax, p[2]);
        }
    Replay rp(sc, walls);
    long mismatches = 0, hits = 0, steps = 0;
    double worst = 0;
    for (int i = 0; i < num_rays; i++) {
        float o[3] = {xmin + (float)urand() * (xmax - xmin), ymin + (float)urand() * (ymax - ymin),
                      zmin + (float)urand() * (zmax - zmin)};
        if (i % 16 == 0) { o[0] = xmin - 3.0f * g.cell - (float)urand(); }            // a few rays start outside the grid
        float d[3];
        double n2;
        do {
            for (int k = 0; k < 3; k++) d[k] = (float)(2.0 * urand() - 1.0);
            n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        } while (n2 < 1e-3 || n2 > 1.0);
        const float inv = 1.0f / sqrtf((float)n2);
        for (int k = 0; k < 3; k++) d[k] *= inv;
        if (i % 97 == 0) d[i % 3] = 0.0f;                                              // axis-parallel component
        if (i % 16 == 0) d[0] = fabsf(d[0]);
        // brute force: strict "<", lowest index wins ties (photonmap.cl:194-206)
        float best = INFINITY;
        int want = -1;
        for (int r = 0; r < cnt[0]; r++) {
            const float t = intersects(walls[r], o, d, best);
            if (t >= 0) { best = t; want = r; }
        }
        float t_got;
        const int got = rp.closest(o, d, t_got, steps);
        hits += want >= 0;
        if (got != want) {
            // a tie between two rectangles at the same distance (shared edges) may resolve either way
            bool tie = false;
            if (got >= 0 && want >= 0) {
                const float t2 = intersects(walls[got], o, d, INFINITY);
                tie = t2 >= 0 && fabsf(t2 - best) <= 1e-4f * fmaxf(best, 1e-3f);
            }
            if (!tie) {
                mismatches++;
                if (mismatches <= 5)
                    printf("mismatch ray %d: o (%g %g %g) d (%g %g %g): brute force %d t %g, table %d t %g\n", i, o[0], o[1], o[2],
                           d[0], d[1], d[2], want, best, got, t_got);
            }
        } else if (want >= 0) {
            worst = std::max(worst, (double)fabsf(t_got - best) / std::max((double)best, 1e-3));
        }
    }
    printf("rays %d hits %ld mismatches %ld worst |dt|/t %.2e steps/ray %.2f\n", num_rays, hits, mismatches, worst,
           (double)steps / num_rays);
    CHECK(hits > num_rays / 10, "too few rays hit anything (%ld)", hits);
    CHECK(mismatches * 100000 <= (long)num_rays, "closest-hit mismatches: %ld of %d", mismatches, num_rays);
    CHECK(worst < 1e-4, "distance error %.2e", worst);
    printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
