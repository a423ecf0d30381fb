// CPU replay of the room tier's photon loop in groups of 32 lanes (tools, not product code): how many warp
// instructions per ray, and how many lanes per instruction, does the loop of k_trace<ROOMS, ..., kRoomSteps> need for a
// given number of box steps per iteration - and what would perfect lanes buy?  The box decomposition is the real one
// (csrc/rooms_build.cpp, the host replay of the device walk), the photons are the bake's in distribution (emitters by
// area, cosine-weighted bounces, depth limit; no floor mirror, no attenuation), the instruction counts per section
// are read off the SASS of the final kernel (profiles/r2i_k_trace_rooms.sass, profiles/r2i_example_sections.txt).
//
//   g++ -O2 -std=c++17 -I include tools/rooms_warp_sim.cpp flatmatch-global-illumination_b200/csrc/rooms_build.cpp \
//       -o /tmp/rooms_warp_sim -lpthread
//   /tmp/rooms_warp_sim scene.bin [depth = 3] [warps = 400] [iterations per warp = 400]
//
// scene.bin: int32 counts (walls, windows, lights) + the three fmgi_rect tables (tests/test_rooms_cpu.py writes it).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../flatmatch-global-illumination_b200/csrc/room_tables.h"

using namespace fmgi;

// warp instructions of one pass through each section (SASS of k_trace<4,0,0,4,0,3>, final build)
static const int kInstRefill = 40;       // A: ballot, popc, chunk bookkeeping, loop head
static const int kInstShade = 125;       // P + S: Philox block, frame loads, sampler, start offset, 1 / d
static const int kInstEmit = 45;         // S, new photons only: second Philox block, emitter loads, start box
static const int kInstCall = 20;         // C: call overhead, step counter, flags
static const int kInstStep = 25;         // one box step
static const int kInstGrid = 18;         // one face-grid lookup
static const int kInstOutcome = 28;      // hit distance (IEEE division), selects
static const int kInstBounce = 100;      // D: hit point, shading record, texel index, roulette, deposit

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline double urand()
{
    rng_state += 0x9E3779B97F4A7C15ull;
    uint64_t z = rng_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

struct Lane {
    bool alive = false, is_new = false, walking = false;
    float o[3], d[3], n[3];
    int box = 0, depth = 0;
};

struct Tally {
    double warp_inst = 0, thread_inst = 0;
    long rays = 0, iterations = 0;
    void add(int inst, int lanes)
    {
        if (lanes <= 0) return;
        warp_inst += inst;
        thread_inst += (double)inst * lanes;
    }
};

static void cosine_direction(const float n[3], float d[3])
{
    // Malley: disk sample lifted to the hemisphere around n
    float u[3] = {n[1], -n[0], 0.0f};
    if (fabsf(n[2]) > 0.9f) { u[0] = 1; u[1] = 0; u[2] = 0; }
    const float ul = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int k = 0; k < 3; k++) u[k] /= ul;
    const float v[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};
    const double r = sqrt(urand()), phi = 6.283184 * urand();
    const float a = (float)(r * cos(phi)), b = (float)(r * sin(phi)), c = (float)sqrt(std::max(0.0, 1.0 - r * r));
    for (int k = 0; k < 3; k++) d[k] = n[k] * c + v[k] * b + u[k] * a;
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s scene.bin [depth] [warps] [iterations]\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("scene"); return 2; }
    int32_t cnt[3];
    if (fread(cnt, 4, 3, f) != 3) return 2;
    std::vector<fmgi_rect> walls(cnt[0]), windows(cnt[1]), lights(cnt[2]);
    if (fread(walls.data(), sizeof(fmgi_rect), cnt[0], f) != (size_t)cnt[0]) return 2;
    if (fread(windows.data(), sizeof(fmgi_rect), cnt[1], f) != (size_t)cnt[1]) return 2;
    if (fread(lights.data(), sizeof(fmgi_rect), cnt[2], f) != (size_t)cnt[2]) return 2;
    fclose(f);
    const int max_depth = argc > 2 ? atoi(argv[2]) : 3, warps = argc > 3 ? atoi(argv[3]) : 400, iters = argc > 4 ? atoi(argv[4]) : 400;
    RoomScene rs;
    const char *why = build_rooms(rs, walls.data(), cnt[0], windows.data(), cnt[1], lights.data(), cnt[2]);
    if (why[0]) { printf("refused: %s\n", why); return 3; }
    std::vector<fmgi_rect> em(windows);
    em.insert(em.end(), lights.begin(), lights.end());
    std::vector<double> cum;
    double tot = 0;
    for (const fmgi_rect &r : em) {
        const double w = sqrt((double)r.width[0] * r.width[0] + r.width[1] * r.width[1] + r.width[2] * r.width[2]);
        const double h = sqrt((double)r.height[0] * r.height[0] + r.height[1] * r.height[1] + r.height[2] * r.height[2]);
        tot += w * h;
        cum.push_back(tot);
    }
    printf("%zu boxes, %zu face grids; depth %d, %d warps x %d iterations\n", rs.boxes.size(), rs.face_grids.size(), max_depth, warps, iters);
    printf("%-10s %14s %12s %12s\n", "steps/iter", "warp-inst/ray", "lanes/inst", "rays/iter");
    double base = 0;
    for (int k_steps : {1, 2, 3, 4, 6, 64}) {
        rng_state = 0x9E3779B97F4A7C15ull;
        Tally t;
        for (int w = 0; w < warps; w++) {
            Lane lane[32];
            for (int it = 0; it < iters; it++) {
                t.iterations++;
                // A. refill
                t.add(kInstRefill, 32);
                for (Lane &L : lane)
                    if (!L.alive) {
                        const double x = urand() * tot;
                        const size_t e = std::min<size_t>(std::lower_bound(cum.begin(), cum.end(), x) - cum.begin(), em.size() - 1);
                        const fmgi_rect &src = em[e];
                        const float a = (float)urand(), b = (float)urand();
                        for (int q = 0; q < 3; q++) { L.o[q] = src.pos[q] + a * src.width[q] + b * src.height[q]; L.n[q] = src.n[q]; }
                        L.alive = true; L.is_new = true; L.walking = false; L.depth = 0;
                        L.box = -2 - (int)e;
                    }
                // P + S for the lanes that are not between boxes
                int shading = 0, emitting = 0;
                for (Lane &L : lane)
                    if (L.alive && !L.walking) {
                        shading++;
                        cosine_direction(L.n, L.d);
                        for (int q = 0; q < 3; q++) L.o[q] += L.d[q] * 1e-5f;
                        if (L.is_new) { emitting++; L.box = rooms_start_box(rs, -2 - L.box, L.o, L.d); L.is_new = false; }
                        t.rays++;
                    }
                t.add(kInstShade, shading);
                t.add(kInstEmit, emitting);
                // C. up to k_steps boxes per lane; the warp runs as many step passes as its slowest lane needs
                t.add(kInstCall, 32);
                int hit_wall[32];
                float hit_t[32];
                bool done[32];
                for (int l = 0; l < 32; l++) { done[l] = !lane[l].alive || lane[l].box < 0; hit_wall[l] = -1; hit_t[l] = 0; if (done[l]) lane[l].alive = false; }
                for (int s = 0; s < k_steps; s++) {
                    int stepping = 0, max_lookups = 0;
                    int lookups_at[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int l = 0; l < 32; l++) {
                        if (done[l]) continue;
                        stepping++;
                        Lane &L = lane[l];
                        // one box: the host replay walks a whole ray; do one step of it by hand
                        const int oct = (L.d[0] > 0 ? 1 : 0) | (L.d[1] > 0 ? 2 : 0) | (L.d[2] > 0 ? 4 : 0);
                        const RoomOctant &R = rs.boxes[(size_t)L.box].oct[oct];
                        float tk[3];
                        for (int q = 0; q < 3; q++) tk[q] = (R.far[q] - L.o[q]) * (L.d[q] == 0 ? -1e30f : 1.0f / L.d[q]);
                        int a = 0;
                        if (tk[1] < tk[0]) a = 1;
                        if (tk[2] < fminf(tk[0], tk[1])) a = 2;
                        const float tt = fminf(fminf(tk[0], tk[1]), tk[2]);
                        const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
                        const float pu = fmaf(tt, L.d[u], L.o[u]), pv = fmaf(tt, L.d[v], L.o[v]);
                        uint32_t code = R.code[a];
                        int lookups = 0;
                        while ((code & kRoomCodeKind) == kRoomCodeNode) {
                            const RoomFaceGrid &G = rs.face_grids[code];
                            const uint32_t iu = (pu >= G.su[0]) + (pu >= G.su[1]) + (pu >= G.su[2]);
                            const uint32_t iv = (pv >= G.sv[0]) + (pv >= G.sv[1]) + (pv >= G.sv[2]);
                            code = rs.face_cells[G.base + iu + G.stride * iv];
                            if (lookups < 8) lookups_at[lookups]++;
                            lookups++;
                        }
                        max_lookups = std::max(max_lookups, lookups);
                        const uint32_t kind = code & kRoomCodeKind, index = code & kRoomCodeIndex;
                        if (kind == kRoomCodeBox) { L.box = (int)index; continue; }
                        done[l] = true;
                        if (kind == kRoomCodeWall && tt >= 0) { hit_wall[l] = (int)index; hit_t[l] = (R.far[a] - L.o[a]) / L.d[a]; }
                        else L.alive = false;
                    }
                    if (!stepping) break;
                    t.add(kInstStep, stepping);
                    for (int q = 0; q < std::min(max_lookups, 8); q++) t.add(kInstGrid, lookups_at[q]);
                }
                // outcome + D for the lanes that hit
                int hits = 0;
                for (int l = 0; l < 32; l++) {
                    Lane &L = lane[l];
                    if (!L.alive) continue;
                    L.walking = !done[l];
                    if (hit_wall[l] < 0) continue;
                    hits++;
                    for (int q = 0; q < 3; q++) { L.o[q] += L.d[q] * hit_t[l]; L.n[q] = walls[(size_t)hit_wall[l]].n[q]; }
                    if (++L.depth == max_depth) L.alive = false;
                }
                t.add(kInstOutcome, hits);
                t.add(kInstBounce, hits);
            }
        }
        const double wi = t.warp_inst / t.rays, lanes = t.thread_inst / t.warp_inst;
        if (k_steps == 3) base = wi;
        printf("%-10d %14.2f %12.2f %12.2f\n", k_steps, wi, lanes, (double)t.rays / t.iterations);
        if (k_steps == 64) {
            printf("\nthread instructions per ray %.0f: with every instruction at 32 lanes %.2f warp instructions per ray, %.2fx the\n"
                   "rate of 3 steps per iteration (%.2f)\n", t.thread_inst / t.rays, t.thread_inst / t.rays / 32.0, base / (t.thread_inst / t.rays / 32.0), base);
        }
    }
    return 0;
}
