// TEST INFRASTRUCTURE (CPU): checks the room tier's box decomposition (csrc/rooms_build.cpp) without a GPU: a host
// replay of the device traversal against a brute-force scan with the reference's intersects() semantics
// (rectangle.c:67-95) on random rays, plus the structure's statistics (leaves, entries, steps and entry tests per ray).
//
// usage: rooms_check scene.bin num_rays
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../flatmatch-global-illumination_b200/csrc/room_tables.h"

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

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.bin num_rays\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("scene"); return 2; }
    int32_t cnt[3];
    if (fread(cnt, 4, 3, f) != 3) return 2;
    std::vector<fmgi_rect> walls(cnt[0]), windows(cnt[1]), lights(cnt[2]);
    if (fread(walls.data(), sizeof(fmgi_rect), cnt[0], f) != (size_t)cnt[0]) return 2;
    if (fread(windows.data(), sizeof(fmgi_rect), cnt[1], f) != (size_t)cnt[1]) return 2;
    if (fread(lights.data(), sizeof(fmgi_rect), cnt[2], f) != (size_t)cnt[2]) return 2;
    fclose(f);
    const int num_rays = atoi(argv[2]);
    RoomScene rs;
    const char *why = build_rooms(rs, walls.data(), cnt[0], windows.data(), cnt[1], lights.data(), cnt[2]);
    if (why[0]) { printf("refused: %s\n", why); return 3; }
    printf("rooms: %zu boxes (%zu kd leaves), %zu face parts (%zu colliders), %zu face grids with %zu cells, depth %d, build %.1f ms\n",
           rs.boxes.size(), rs.kd_leaves, rs.face_parts, rs.wall_parts, rs.face_grids.size(), rs.face_cells.size(), rs.max_depth, rs.build_ms);

    // a checksum of every table the device reads: the build must not depend on the number of builder threads
    {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const void *data, size_t bytes) {
            const unsigned char *b = static_cast<const unsigned char *>(data);
            for (size_t i = 0; i < bytes; i++) { h ^= b[i]; h *= 1099511628211ull; }
        };
        mix(rs.boxes.data(), rs.boxes.size() * sizeof(RoomBox));
        mix(rs.bounds.data(), rs.bounds.size() * sizeof(RoomBounds));
        mix(rs.face_grids.data(), rs.face_grids.size() * sizeof(RoomFaceGrid));
        mix(rs.face_cells.data(), rs.face_cells.size() * sizeof(uint32_t));
        mix(rs.nodes.data(), rs.nodes.size() * sizeof(RoomNode));
        mix(rs.starts.data(), rs.starts.size() * sizeof(RoomStart));
        printf("tables checksum %016llx\n", (unsigned long long)h);
    }

    // rays as the path produces them: half start inside the bounding box, half on a wall (offset 1e-5 along the ray)
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (const fmgi_rect &r : walls)
        for (int c = 0; c < 4; c++)
            for (int k = 0; k < 3; k++) {
                const float x = r.pos[k] + (c & 1 ? r.width[k] : 0.0f) + (c & 2 ? r.height[k] : 0.0f);
                lo[k] = fminf(lo[k], x); hi[k] = fmaxf(hi[k], x);
            }
    long steps = 0, tests = 0, mism = 0, mism_edge = 0, hits = 0, max_steps = 0;
    bool bad_starts = false;
    std::vector<long> hist(32, 0), ehist(32, 0);
    for (int i = 0; i < num_rays; i++) {
        float o[3], d[3];
        double n2 = 0;
        for (int k = 0; k < 3; k++) { d[k] = (float)(urand() * 2 - 1); n2 += (double)d[k] * d[k]; }
        if (n2 < 1e-4 || n2 > 1) { i--; continue; }
        for (int k = 0; k < 3; k++) d[k] = (float)(d[k] / sqrt(n2));
        if (i & 1) {
            const fmgi_rect &r = walls[(size_t)(urand() * walls.size()) % walls.size()];
            const float a = (float)urand(), b = (float)urand();
            float dn = 0;
            for (int k = 0; k < 3; k++) dn += d[k] * r.n[k];
            if (dn < 0) for (int k = 0; k < 3; k++) d[k] -= 2 * dn * r.n[k];
            for (int k = 0; k < 3; k++) o[k] = r.pos[k] + a * r.width[k] + b * r.height[k] + d[k] * 1e-5f;
        } else {
            for (int k = 0; k < 3; k++) o[k] = (float)(lo[k] + urand() * (hi[k] - lo[k]));
        }
        // brute force: strict "<", lowest index wins ties
        float best = INFINITY;
        int want = -1;
        for (size_t w = 0; w < walls.size(); w++) {
            const float t = intersects(walls[w], o, d, best);
            if (t >= 0) { best = t; want = (int)w; }
        }
        float t;
        int leaf_out;
        long s0 = steps, e0 = tests;
        const int leaf = rooms_locate(rs, o, d);
        const int got = leaf < 0 ? -1 : rooms_closest_hit(rs, leaf, o, d, t, leaf_out, steps, tests);
        const long st = steps - s0;
        max_steps = std::max(max_steps, st);
        hist[std::min<long>(st, 31)]++;
        ehist[std::min<long>(tests - e0, 31)]++;
        hits += got >= 0;
        if (got != want) {
            mism++;
            if (mism <= 5)
                printf("mismatch ray %d: got %d want %d (t %.7g vs %.7g) o=(%.6f %.6f %.6f) d=(%.6f %.6f %.6f)\n", i, got, want,
                       got >= 0 ? t : -1.0f, best, o[0], o[1], o[2], d[0], d[1], d[2]);
        } else if (got >= 0 && fabsf(t - best) > 1e-4f * fmaxf(best, 1e-3f)) {
            mism_edge++;
        }
    }
    printf("rays %d: hits %.3f, steps/ray %.3f (max %ld), face lookups/ray %.3f, index mismatches %ld, distance mismatches %ld\n",
           num_rays, (double)hits / num_rays, (double)steps / num_rays, max_steps, (double)tests / num_rays, mism, mism_edge);
    printf("steps histogram:");
    for (int i = 0; i < 16; i++) printf(" %d:%.3f", i, (double)hist[i] / num_rays);
    // the rays a bake traces: photons from the emitters (by area), diffuse bounces, up to 4 deposits
    {
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
        long psteps = 0, pnodes = 0, prays = 0, start_checks = 0, start_mismatch = 0;
        std::vector<long> ph(32, 0), pn(32, 0);
        for (int i = 0; i < num_rays / 2 && !em.empty(); i++) {
            const double x = urand() * tot;
            const size_t e = std::lower_bound(cum.begin(), cum.end(), x) - cum.begin();
            const fmgi_rect *src = &em[std::min(e, em.size() - 1)];
            float o[3], d[3], nrm[3] = {src->n[0], src->n[1], src->n[2]};
            const float a = (float)urand(), b = (float)urand();
            for (int k = 0; k < 3; k++) o[k] = src->pos[k] + a * src->width[k] + b * src->height[k];
            int box = -2;
            for (int depth = 0; depth < 4; depth++) {
                double n2;
                do {
                    n2 = 0;
                    for (int k = 0; k < 3; k++) { d[k] = (float)(urand() * 2 - 1); n2 += (double)d[k] * d[k]; }
                } while (n2 < 1e-4 || n2 > 1);
                float dn = 0;
                for (int k = 0; k < 3; k++) { d[k] = (float)(d[k] / sqrt(n2)); dn += d[k] * nrm[k]; }
                if (dn < 0) for (int k = 0; k < 3; k++) d[k] -= 2 * dn * nrm[k];
                for (int k = 0; k < 3; k++) o[k] += d[k] * 1e-5f;
                if (box == -2) {
                    box = rooms_locate(rs, o, d);
                    // the emitter's partition must name the box the tree descent finds (device: rooms_start)
                    const int emitter = (int)std::min(e, em.size() - 1);
                    start_checks++;
                    if (rooms_start_box(rs, emitter, o, d) != box) {
                        // a start point within 1e-5 of a box edge may sit in the neighbour: count, tolerate a few
                        start_mismatch++;
                    }
                }
                if (box < 0) break;
                long s0 = steps, n0 = tests;
                float t;
                int box_out;
                const int got = rooms_closest_hit(rs, box, o, d, t, box_out, steps, tests);
                prays++;
                psteps += steps - s0; pnodes += tests - n0;
                ph[std::min<long>(steps - s0, 31)]++; pn[std::min<long>(tests - n0, 31)]++;
                steps = s0; tests = n0;
                if (got < 0) break;
                for (int k = 0; k < 3; k++) { o[k] += d[k] * t; nrm[k] = walls[got].n[k]; }
                box = box_out;
            }
        }
        printf("\nstart boxes: %ld photons, %ld not in the box the tree descent finds", start_checks, start_mismatch);
        if (start_mismatch > start_checks / 2000 + 1) bad_starts = true;
        if (prays) {
            printf("\nphoton rays %ld: steps/ray %.3f, face lookups/ray %.3f\nphoton steps histogram:", prays, (double)psteps / prays,
                   (double)pnodes / prays);
            for (int i = 0; i < 12; i++) printf(" %d:%.3f", i, (double)ph[i] / prays);
            printf("\nphoton face lookups histogram:");
            for (int i = 0; i < 12; i++) printf(" %d:%.3f", i, (double)pn[i] / prays);
        }
    }
    printf("\nface lookups histogram:");
    for (int i = 0; i < 16; i++) printf(" %d:%.3f", i, (double)ehist[i] / num_rays);
    printf("\n");
    return mism > num_rays / 20000 + 2 || mism_edge || bad_starts ? 1 : 0;
}
