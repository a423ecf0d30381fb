// Design-exploration tool (host only, not part of the product or of the tests): replays the photon
// loop of k_trace<GRID> on the CPU in groups of 32 lanes and counts, per ray, how many warp
// iterations / issue slots different shapes of the grid-walk loop would need.  GPU minutes are scarce;
// this answers "which loop structure and which cell size" before a kernel is written.
//
//   g++ -O2 -std=c++17 -ffp-contract=off -I include tools/simt_walk_sim.cpp \
//       flatmatch-global-illumination_b200/csrc/scene_prep.cpp -o /tmp/simt_walk_sim
//   /tmp/simt_walk_sim scene.bin [cell] [photon warps] [depth]
// scene.bin: int32 numWalls, numWindows, numLights, then the three fmgi_rect tables (tools/dump_scene.py).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../flatmatch-global-illumination_b200/csrc/scene_tables.h"

using namespace fmgi;
struct int2 { int x, y; };

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline float frand()
{
    rng_state += 0x9E3779B97F4A7C15ull;
    uint64_t z = rng_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

struct Walk {                       // what one closest-hit query did
    int plane_heads = 0;            // planes whose crossing point was looked up
    int plane_tests = 0;
    std::vector<int> plane_cands;   // candidates tested per looked-up plane
    std::vector<int> cells;         // records in the walk list of every visited cell
    int hit = -1;
    float t = INFINITY;
};

struct Sim {
    HostScene sc;
    std::vector<fmgi_rect> walls, windows, lights;

    // the kernel's order (GridWalk in csrc/trace_kernels.cuh): walls by the 2-D DDA, bounded by the grid box
    // and the z range of the walls, then the z planes crossed before the wall hit, nearest first
    int closest(const float o[3], const float d[3], Walk &w) const
    {
        const GridDesc &g = sc.grid;
        const int ncell = g.nx * g.ny;
        const int2 *ranges = reinterpret_cast<const int2 *>(sc.grid_ranges.data());
        float best = INFINITY;
        int win = -1;
        const float ix = 1.0f / d[0], iy = 1.0f / d[1];
        int cx = (int)floorf((o[0] - g.x0) * g.inv_cell), cy = (int)floorf((o[1] - g.y0) * g.inv_cell);
        cx = std::min(std::max(cx, 0), g.nx - 1); cy = std::min(std::max(cy, 0), g.ny - 1);
        float tmx = INFINITY, tmy = INFINITY, tdx = INFINITY, tdy = INFINITY, t_exit = INFINITY;
        if (d[0] != 0) {
            tmx = (g.x0 + (float)(cx + (d[0] > 0 ? 1 : 0)) * g.cell - o[0]) * ix;
            tdx = g.cell * fabsf(ix);
            t_exit = ((d[0] > 0 ? g.exit_hi_x : g.exit_lo_x) - o[0]) * ix;
        }
        if (d[1] != 0) {
            tmy = (g.y0 + (float)(cy + (d[1] > 0 ? 1 : 0)) * g.cell - o[1]) * iy;
            tdy = g.cell * fabsf(iy);
            t_exit = fminf(t_exit, ((d[1] > 0 ? g.exit_hi_y : g.exit_lo_y) - o[1]) * iy);
        }
        if (d[2] != 0) t_exit = fminf(t_exit, ((d[2] < 0 ? g.wall_z_lo : g.wall_z_hi) - o[2]) / d[2] * 1.0001f);
        const int sx = d[0] > 0 ? 1 : -1, sy = d[1] > 0 ? g.nx : -g.nx;
        int ci = cy * g.nx + cx;
        const int combo = (d[0] > 0 ? 1 : 0) + (d[1] > 0 ? 2 : 0);
        const int2 *walk = ranges + (kWalkListBase + combo) * ncell;
        best = fminf(best, t_exit);
        for (;;) {
            const int2 r = walk[ci];
            w.cells.push_back(r.y - r.x);
            for (int q = r.x; q < r.y; q++) {
                const GridRec &rec = sc.grid_recs[q];
                if (rec.tag & kTagMisc) continue;      // parseLayout scenes have none in walk lists
                const bool ky = (rec.tag & kTagAlongY) != 0;
                const float t = (rec.c - (ky ? o[1] : o[0])) * (ky ? iy : ix);
                const float pi = t * (ky ? d[0] : d[1]) + (ky ? o[0] : o[1]) - rec.mid_i;
                const float pj = t * d[2] + o[2] - rec.mid_j;
                if (t >= 0 && t < best && fabsf(pi) <= rec.half_i && fabsf(pj) <= rec.half_j) { best = t; win = q; }
            }
            const float t_next = fminf(tmx, tmy);
            if (!(t_next < best)) break;
            if (tmx < tmy) { ci += sx; tmx += tdx; } else { ci += sy; tmy += tdy; }
        }
        if (win < 0) best = INFINITY;
        if (d[2] != 0.0f) {
            const float iz = 1.0f / d[2];
            const int first = d[2] < 0 ? 0 : kMaxPlanesPerSign;
            const int count = d[2] < 0 ? g.planes_up : g.planes_down;
            for (int pl = 0; pl < count; pl++) {
                const float t = (g.plane_z[first + pl] - o[2]) * iz;
                if (!(t >= 0 && t < best)) continue;
                const float x = t * d[0] + o[0], y = t * d[1] + o[1];
                const int px = (int)floorf((x - g.x0) * g.inv_cell), py = (int)floorf((y - g.y0) * g.inv_cell);
                if (px < 0 || py < 0 || px >= g.nx || py >= g.ny) continue;
                w.plane_heads++;
                const int2 r = ranges[(first + pl) * ncell + py * g.nx + px];
                int c = 0;
                for (int q = r.x; q < r.y; q++) {
                    const GridRec &rec = sc.grid_recs[q];
                    c++;
                    w.plane_tests++;
                    if (fabsf(x - rec.mid_i) <= rec.half_i && fabsf(y - rec.mid_j) <= rec.half_j) { best = t; win = q; break; }
                }
                w.plane_cands.push_back(c);
            }
        }
        w.t = best;
        w.hit = win >= 0 ? (int)(sc.grid_recs[win].tag & kTagIdMask) : -1;
        return w.hit;
    }
};

static void sample(const float n[3], const float u[3], const float v[3], bool sky, float out[3])
{
    const float r = sqrtf(frand()), phi = 2.0f * 3.141592f * frand();
    float a = r * cosf(phi);
    const float b = r * sinf(phi), c = sqrtf(1.0f - r * r);
    if (sky) a = fabsf(a);
    for (int k = 0; k < 3; k++) out[k] = n[k] * c + v[k] * b + u[k] * a;
}

struct Lane {
    bool alive = false, mirror = false, is_new = false;
    float p[3], d[3];
    int depth = 0, hit = 0, emitter = 0;
};

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s scene.bin [cell] [warps] [depth]\n", argv[0]); return 1; }
    const float cell = argc > 2 ? (float)atof(argv[2]) : 0.0f;
    const int num_warps = argc > 3 ? atoi(argv[3]) : 2000;
    const int max_depth = argc > 4 ? atoi(argv[4]) : 4;
    const int photons_per_lane = 12;
    Sim sim;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("scene"); return 1; }
    int32_t cnt[3];
    if (fread(cnt, 4, 3, f) != 3) return 1;
    sim.walls.resize(cnt[0]); sim.windows.resize(cnt[1]); sim.lights.resize(cnt[2]);
    if (fread(sim.walls.data(), sizeof(fmgi_rect), cnt[0], f) != (size_t)cnt[0]) return 1;
    if (fread(sim.windows.data(), sizeof(fmgi_rect), cnt[1], f) != (size_t)cnt[1]) return 1;
    if (fread(sim.lights.data(), sizeof(fmgi_rect), cnt[2], f) != (size_t)cnt[2]) return 1;
    fclose(f);
    const char *why = prepare_scene(sim.sc, sim.walls.data(), cnt[0], sim.windows.data(), cnt[1], sim.lights.data(), cnt[2], 1 << 30);
    if (why[0] && strstr(why, "atlas") == nullptr) { fprintf(stderr, "scene: %s\n", why); }
    build_grid(sim.sc, sim.walls.data(), cnt[0], sim.windows.data(), cnt[1], sim.lights.data(), cnt[2], cell);
    const GridDesc &g = sim.sc.grid;
    printf("grid %dx%d cell %.3f recs %zu planes %d/%d\n", g.nx, g.ny, g.cell, sim.sc.grid_recs.size(), g.planes_up, g.planes_down);

    // emitter CDF by area
    std::vector<double> cdf;
    double tot = 0;
    for (float a : sim.sc.emitter_area) { tot += a; cdf.push_back(tot); }

    // accumulators
    double rays = 0, tests = 0, cells_visited = 0, empty_cells = 0;
    double v0_test_iters = 0, v0_adv_iters = 0, v0_iters = 0;       // current kernel: warp iterations in which any lane tests / advances
    double v0_test_lanes = 0, v0_adv_lanes = 0;
    double v1_iters = 0, v1_lanes = 0;                              // merged: one iteration = [test] + [advance if exhausted]
    double v2_iters = 0, v2_lanes = 0;                              // merged, two tests per iteration
    double pl_head_iters = 0, pl_cand_iters = 0, pl_head_lanes = 0, pl_cand_lanes = 0;
    double outer = 0, lanes_alive = 0;
    std::vector<double> hist(64, 0.0), hist_max(64, 0.0);

    for (int wq = 0; wq < num_warps; wq++) {
        Lane L[32];
        int budget[32];
        for (int l = 0; l < 32; l++) budget[l] = photons_per_lane;
        for (;;) {
            // refill
            for (int l = 0; l < 32; l++)
                if (!L[l].alive && budget[l] > 0) {
                    budget[l]--;
                    L[l].alive = true; L[l].is_new = true; L[l].depth = 0;
                    const double x = frand() * tot;
                    L[l].emitter = (int)(std::lower_bound(cdf.begin(), cdf.end(), x) - cdf.begin());
                    if (L[l].emitter >= (int)cdf.size()) L[l].emitter = (int)cdf.size() - 1;
                }
            int alive = 0;
            for (int l = 0; l < 32; l++) alive += L[l].alive;
            if (!alive) break;
            outer++; lanes_alive += alive;
            Walk W[32];
            for (int l = 0; l < 32; l++) {
                Lane &a = L[l];
                if (!a.alive) continue;
                if (a.is_new) {
                    const EmitterRec &e = sim.sc.emitters[a.emitter];
                    sample(e.n, e.u, e.v, e.is_window != 0, a.d);
                    const float sx = frand(), sy = frand();
                    for (int k = 0; k < 3; k++) a.p[k] = e.pos[k] + a.d[k] * 1e-5f + e.width[k] * sx + e.height[k] * sy;
                    a.is_new = false;
                } else {
                    const ShadeRect &s = sim.sc.shade[a.hit];
                    if (a.mirror) {
                        const float k2 = 2.0f * (s.n[0] * a.d[0] + s.n[1] * a.d[1] + s.n[2] * a.d[2]);
                        for (int k = 0; k < 3; k++) a.d[k] -= k2 * s.n[k];
                    } else {
                        sample(s.n, s.u, s.v, false, a.d);
                    }
                    for (int k = 0; k < 3; k++) a.p[k] += a.d[k] * 1e-5f;
                }
                sim.closest(a.p, a.d, W[l]);
                rays++;
            }
            // ---- cost models over the 32 walks -------------------------------------------------
            // plane phase (same in all variants): loop over planes, inner candidate loop
            {
                size_t maxh = 0;
                for (int l = 0; l < 32; l++) if (L[l].alive) maxh = std::max(maxh, W[l].plane_cands.size());
                for (size_t h = 0; h < maxh; h++) {
                    int lanes = 0, maxc = 0;
                    for (int l = 0; l < 32; l++)
                        if (L[l].alive && h < W[l].plane_cands.size()) { lanes++; maxc = std::max(maxc, W[l].plane_cands[h]); }
                    pl_head_iters++; pl_head_lanes += lanes;
                    for (int c = 0; c < maxc; c++) {
                        int ll = 0;
                        for (int l = 0; l < 32; l++)
                            if (L[l].alive && h < W[l].plane_cands.size() && c < W[l].plane_cands[h]) ll++;
                        pl_cand_iters++; pl_cand_lanes += ll;
                    }
                }
            }
            // V0: op sequence per lane: T*n0, A, T*n1, A, ..., last A (fails)
            // V1: per cell max(n,1) iterations;  V2: per cell max(ceil(n/2),1)
            {
                std::vector<char> ops[32];
                int len1[32], len2[32];
                size_t max0 = 0; int max1 = 0, max2 = 0;
                for (int l = 0; l < 32; l++) {
                    len1[l] = len2[l] = 0;
                    if (!L[l].alive) continue;
                    for (int n : W[l].cells) {
                        for (int q = 0; q < n; q++) ops[l].push_back('T');
                        ops[l].push_back('A');
                        len1[l] += std::max(n, 1);
                        len2[l] += std::max((n + 1) / 2, 1);
                        tests += n; cells_visited++; empty_cells += n == 0;
                    }
                    tests += W[l].plane_tests;
                    max0 = std::max(max0, ops[l].size()); max1 = std::max(max1, len1[l]); max2 = std::max(max2, len2[l]);
                    hist[std::min(len1[l], 63)]++;
                }
                for (size_t j = 0; j < max0; j++) {
                    int t = 0, a = 0;
                    for (int l = 0; l < 32; l++) if (j < ops[l].size()) { if (ops[l][j] == 'T') t++; else a++; }
                    v0_iters++;
                    if (t) { v0_test_iters++; v0_test_lanes += t; }
                    if (a) { v0_adv_iters++; v0_adv_lanes += a; }
                }
                v1_iters += max1; v2_iters += max2;
                hist_max[std::min(max1, 63)]++;
                for (int l = 0; l < 32; l++) { v1_lanes += len1[l]; v2_lanes += len2[l]; }
            }
            // ---- shade --------------------------------------------------------------------------
            for (int l = 0; l < 32; l++) {
                Lane &a = L[l];
                if (!a.alive) continue;
                if (W[l].hit < 0) { a.alive = false; continue; }
                a.hit = W[l].hit;
                for (int k = 0; k < 3; k++) a.p[k] += a.d[k] * W[l].t;
                a.mirror = a.p[2] < 0.0005f && frand() < 0.75f;
                a.depth++;
                if (a.depth == max_depth) a.alive = false;
            }
        }
    }
    printf("rays %.0f  tests/ray %.2f  cells/ray %.2f (empty %.2f)  alive lanes/outer %.1f\n", rays, tests / rays,
           cells_visited / rays, empty_cells / rays, lanes_alive / outer);
    printf("planes: head iters/ray %.3f (lanes %.1f)  cand iters/ray %.3f (lanes %.1f)\n", pl_head_iters / rays,
           pl_head_lanes / pl_head_iters, pl_cand_iters / rays, pl_cand_lanes / pl_cand_iters);
    printf("V0 current : warp iters/outer %.2f  test iters %.2f (lanes %.1f)  adv iters %.2f (lanes %.1f)\n", v0_iters / outer,
           v0_test_iters / outer, v0_test_lanes / v0_test_iters, v0_adv_iters / outer, v0_adv_lanes / v0_adv_iters);
    printf("V1 merged  : warp iters/outer %.2f  lanes active %.1f   lane iters/ray %.2f\n", v1_iters / outer, v1_lanes / v1_iters, v1_lanes / rays);
    printf("V2 merged x2: warp iters/outer %.2f  lanes active %.1f   lane iters/ray %.2f\n", v2_iters / outer, v2_lanes / v2_iters, v2_lanes / rays);
    // issue-slot estimates per outer iteration (instruction counts are guesses)
    const double cT = 20, cA = 14, cL = 4, cM = 30, cM2 = 42;
    printf("est. walk issue slots/outer: V0 %.0f  V1 %.0f  V2 %.0f\n",
           (v0_iters * cL + v0_test_iters * cT + v0_adv_iters * cA) / outer, v1_iters * cM / outer, v2_iters * cM2 / outer);
    printf("V1 length histogram:");
    double hs = 0; for (double h : hist) hs += h;
    for (int i = 0; i < 24; i++) printf(" %d:%.3f", i, hist[i] / hs);
    printf("\n");
    // the same per warp: steps of the warp's longest walk = iterations the whole warp spends in the walk loop
    printf("warp-max histogram:");
    double hm = 0, mean_max = 0; for (double h : hist_max) hm += h;
    for (int i = 0; i < 64; i++) mean_max += i * hist_max[i] / hm;
    for (int i = 0; i < 40; i++) printf(" %d:%.3f", i, hist_max[i] / hm);
    printf("\nmean steps per ray %.2f, mean of the warp maximum %.2f\n", v1_lanes / rays, mean_max);
    // ---- model R: walk capped at M iterations per round; unfinished lanes keep walking next round and
    //      skip the shade/emit phase ("if-if" scheduling).  Model P: per-warp pool of 32*K rays, lanes
    //      fetch the next ray when at least `thr` lanes are idle.
    {
        const double cS = argc > 5 ? atof(argv[5]) : 400, cMm = argc > 6 ? atof(argv[6]) : 35, cF = 14;
        auto run_photons = [&](auto &&next_len) { (void)next_len; };
        (void)run_photons;
        for (int M : {4, 6, 8, 10, 12, 16, 1000}) {
            rng_state = 12345;
            double nrays = 0, rounds = 0, iters = 0, s_lanes = 0;
            for (int wq = 0; wq < num_warps / 4; wq++) {
                Lane L[32];
                int budget[32], remaining[32];
                Walk W[32];
                for (int l = 0; l < 32; l++) { budget[l] = photons_per_lane; remaining[l] = 0; }
                for (;;) {
                    int need = 0;
                    // S phase for lanes whose walk is over
                    for (int l = 0; l < 32; l++) {
                        Lane &a = L[l];
                        if (remaining[l] > 0) continue;
                        if (a.alive && !a.is_new && W[l].cells.size()) {     // shade the finished walk
                            if (W[l].hit < 0) a.alive = false;
                            else {
                                a.hit = W[l].hit;
                                for (int k = 0; k < 3; k++) a.p[k] += a.d[k] * W[l].t;
                                a.mirror = a.p[2] < 0.0005f && frand() < 0.75f;
                                if (++a.depth == max_depth) a.alive = false;
                            }
                        }
                        if (!a.alive && budget[l] > 0) {
                            budget[l]--;
                            a.alive = true; a.is_new = true; a.depth = 0;
                            const double x = frand() * tot;
                            a.emitter = std::min((int)(std::lower_bound(cdf.begin(), cdf.end(), x) - cdf.begin()), (int)cdf.size() - 1);
                        }
                        if (!a.alive) { W[l] = Walk(); continue; }
                        if (a.is_new) {
                            const EmitterRec &e = sim.sc.emitters[a.emitter];
                            sample(e.n, e.u, e.v, e.is_window != 0, a.d);
                            const float sx = frand(), sy = frand();
                            for (int k = 0; k < 3; k++) a.p[k] = e.pos[k] + a.d[k] * 1e-5f + e.width[k] * sx + e.height[k] * sy;
                            a.is_new = false;
                        } else {
                            const ShadeRect &s = sim.sc.shade[a.hit];
                            if (a.mirror) {
                                const float k2 = 2.0f * (s.n[0] * a.d[0] + s.n[1] * a.d[1] + s.n[2] * a.d[2]);
                                for (int k = 0; k < 3; k++) a.d[k] -= k2 * s.n[k];
                            } else sample(s.n, s.u, s.v, false, a.d);
                            for (int k = 0; k < 3; k++) a.p[k] += a.d[k] * 1e-5f;
                        }
                        W[l] = Walk();
                        sim.closest(a.p, a.d, W[l]);
                        int len = 0;
                        for (int n : W[l].cells) len += std::max(n, 1);
                        remaining[l] = len;
                        nrays++; need++;
                    }
                    int mx = 0;
                    for (int l = 0; l < 32; l++) mx = std::max(mx, remaining[l]);
                    if (!mx) break;
                    rounds++; s_lanes += need;
                    const int it = std::min(mx, M);
                    iters += it;
                    for (int l = 0; l < 32; l++) remaining[l] -= std::min(remaining[l], it);
                }
            }
            printf("R M=%4d: rounds/ray %.4f (S lanes %.1f)  walk iters/ray %.3f  est slots/ray %.1f\n", M, rounds / nrays,
                   s_lanes / rounds, iters / nrays, (rounds * cS + iters * cMm) / nrays);
        }
        (void)cF;
    }
    // ---- models Q / P: a warp walks 32 * K rays per round.  Q: every lane owns a private queue of K rays
    //      (static); P: idle lanes take the next ray of the warp's shared pool (dynamic).  Idle lanes
    //      switch to their next ray together, when at least `thr` lanes are idle (or nobody is walking).
    {
        // ray walk lengths of the scene, in emission order of a long run (recorded by a plain replay)
        std::vector<int> lens;
        double miss_len = 0, miss_n = 0, hit_len = 0, hit_n = 0, miss_long = 0, hit_long = 0;
        {
            rng_state = 777;
            for (int wq = 0; wq < num_warps / 2; wq++) {
                Lane L[32];
                int budget[32];
                for (int l = 0; l < 32; l++) budget[l] = photons_per_lane;
                for (;;) {
                    int alive = 0;
                    for (int l = 0; l < 32; l++) {
                        Lane &a = L[l];
                        if (!a.alive && budget[l] > 0) {
                            budget[l]--;
                            a.alive = true; a.is_new = true; a.depth = 0;
                            const double x = frand() * tot;
                            a.emitter = std::min((int)(std::lower_bound(cdf.begin(), cdf.end(), x) - cdf.begin()), (int)cdf.size() - 1);
                        }
                        alive += a.alive;
                    }
                    if (!alive) break;
                    for (int l = 0; l < 32; l++) {
                        Lane &a = L[l];
                        if (!a.alive) continue;
                        if (a.is_new) {
                            const EmitterRec &e = sim.sc.emitters[a.emitter];
                            sample(e.n, e.u, e.v, e.is_window != 0, a.d);
                            const float sx = frand(), sy = frand();
                            for (int k = 0; k < 3; k++) a.p[k] = e.pos[k] + a.d[k] * 1e-5f + e.width[k] * sx + e.height[k] * sy;
                            a.is_new = false;
                        } else {
                            const ShadeRect &sh = sim.sc.shade[a.hit];
                            if (a.mirror) {
                                const float k2 = 2.0f * (sh.n[0] * a.d[0] + sh.n[1] * a.d[1] + sh.n[2] * a.d[2]);
                                for (int k = 0; k < 3; k++) a.d[k] -= k2 * sh.n[k];
                            } else sample(sh.n, sh.u, sh.v, false, a.d);
                            for (int k = 0; k < 3; k++) a.p[k] += a.d[k] * 1e-5f;
                        }
                        Walk w;
                        sim.closest(a.p, a.d, w);
                        int len = 0;
                        for (int n : w.cells) len += std::max(n, 1);
                        lens.push_back(len);
                        if (w.hit < 0) { miss_len += len; miss_n++; if (len > 8) miss_long++; } else { hit_len += len; hit_n++; if (len > 8) hit_long++; }
                        if (w.hit < 0) { a.alive = false; continue; }
                        a.hit = w.hit;
                        for (int k = 0; k < 3; k++) a.p[k] += a.d[k] * w.t;
                        a.mirror = a.p[2] < 0.0005f && frand() < 0.75f;
                        if (++a.depth == max_depth) a.alive = false;
                    }
                }
            }
        }
        printf("walk length: hits mean %.2f (n %.0f, >8: %.3f)  misses mean %.2f (n %.0f, >8: %.3f)\n", hit_len / hit_n, hit_n, hit_long / hit_n, miss_len / std::max(miss_n, 1.0), miss_n, miss_long / std::max(miss_n, 1.0));
        { double m = 0; int mx = 0; for (int v : lens) { m += v; mx = std::max(mx, v); } printf("lens: n %zu mean %.2f max %d\n", lens.size(), m / lens.size(), mx); }
        for (int dynamic = 0; dynamic < 2; dynamic++)
            for (int K : {1, 2, 4, 8})
                for (int thr : {1, 8, 16, 24}) {
                    double iters = 0, switches = 0, sw_lanes = 0, nr = 0, act = 0;
                    size_t pos = 0;
                    while (pos + 32 * K <= lens.size()) {
                        const int *pool = &lens[pos];
                        pos += 32 * K;
                        nr += 32 * K;
                        int rem[32], next_own[32], pool_next = 0;
                        for (int l = 0; l < 32; l++) { rem[l] = 0; next_own[l] = 0; }
                        for (;;) {
                            int idle_with_work = 0, walking = 0;
                            for (int l = 0; l < 32; l++) {
                                if (rem[l] > 0) walking++;
                                else if (dynamic ? pool_next < 32 * K : next_own[l] < K) idle_with_work++;
                            }
                            if (!walking && !idle_with_work) break;
                            if (idle_with_work && (idle_with_work >= thr || !walking)) {
                                switches++;
                                for (int l = 0; l < 32; l++)
                                    if (rem[l] == 0) {
                                        if (dynamic) { if (pool_next < 32 * K) { rem[l] = pool[pool_next++]; sw_lanes++; } }
                                        else if (next_own[l] < K) { rem[l] = pool[next_own[l]++ * 32 + l]; sw_lanes++; }
                                    }
                                continue;
                            }
                            iters++;
                            for (int l = 0; l < 32; l++) if (rem[l] > 0) { rem[l]--; act++; }
                        }
                    }
                    printf("%s K=%d thr=%2d: warp iters/ray %.3f (lanes %.1f)  switch events/ray %.4f (lanes %.1f)\n", dynamic ? "P" : "Q", K, thr,
                           iters / nr, act / iters, switches / nr, sw_lanes / switches);
                }
    }
    return 0;
}
