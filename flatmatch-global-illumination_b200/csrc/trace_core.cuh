// The persistent photon-tracing kernel (grid tier: floor-plan grid walked from L1 / L2; soup tier: whole
// rectangle soup in shared memory) plus the probe kernels that expose its device functions to the
// parity tests.
#pragma once
#include "trace_kernels.cuh"

namespace fmgi {

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// Room tier for a ray that may start anywhere (probes, ambient occlusion): a ray from outside the root box enters
// it where the slab test says (or misses it), then the first box is located by tree descent.
__device__ __forceinline__ int closest_hit_rooms_from_anywhere(const TraceParams &p, float ox, float oy, float oz, float dx,
                                                               float dy, float dz, float &t_out)
{
    const float inf = __int_as_float(0x7f800000);
    t_out = inf;
    float t0 = 0.0f, t1 = inf;
    const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (d[k] == 0.0f) {
            if (o[k] < p.room_lo[k] || o[k] > p.room_hi[k]) return -1;
        } else {
            const float ta = (p.room_lo[k] - o[k]) / d[k], tb = (p.room_hi[k] - o[k]) / d[k];
            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
        }
    }
    if (!(t0 <= t1)) return -1;                                // the ray misses the scene's box
    int leaf = rooms_locate(p, fmaf(t0, dx, ox), fmaf(t0, dy, oy), fmaf(t0, dz, oz), dx, dy, dz);
    return closest_hit_rooms(p, leaf, ox, oy, oz, dx, dy, dz, t_out);
}

// Binary search: largest e with job_begin[e] <= job (job < total_jobs).
__device__ __forceinline__ int find_emitter(const unsigned long long *__restrict__ job_begin, int num_emitters,
                                            unsigned long long job)
{
    int lo = 0, hi = num_emitters - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(job_begin + mid) <= job) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Stages the soup into shared memory (whole CTA) and returns the table view.
__device__ __forceinline__ SoupTables stage_soup(const TraceParams &p, float4 *smem)
{
    const int n_axis4 = 6 * p.pair_begin[3];
    const int n_gen4 = 4 * p.num_general;
    for (int i = threadIdx.x; i < n_axis4; i += blockDim.x) smem[i] = p.axis[i];
    for (int i = threadIdx.x; i < n_gen4; i += blockDim.x) smem[n_axis4 + i] = p.general[i];
    __syncthreads();
    SoupTables soup;
    soup.axis = smem;
    soup.general = smem + n_axis4;
#pragma unroll
    for (int g = 0; g < 4; g++) soup.pair_begin[g] = p.pair_begin[g];
    soup.num_general = p.num_general;
    return soup;
}

// kTier: FMGI_TIER_SOUP (brute force over the shared-memory soup), kTierSoupPlanes (the same with the
// horizontal rectangles looked up through the grid's plane tables), FMGI_TIER_GRID (floor-plan grid in L2) or
// kTierRooms (box decomposition, the default for axis-parallel scenes).
// kCount: also count the rectangle tests the grid lookups execute (fmgi_options.count_tests; two more
// instructions in the walk loop, so not the default).
// kRoomSteps (room tier): boxes a lane's ray crosses per iteration of the photon loop.  The room tier does not walk a
// ray to its hit before the warp goes on: a lane whose ray is still between boxes after kRoomSteps keeps walking in
// the next iteration while the lanes that hit something shade, deposit and draw their next ray - the warp never
// waits for its longest walk.
template <int kTier, int kDeposit, bool kProbe, int kMinBlocks, bool kCount = false, int kRoomSteps = 2>
__global__ void __launch_bounds__(kTraceThreads, kMinBlocks) k_trace(const TraceParams p)
{
    extern __shared__ float4 smem[];
    SoupTables soup;
    if (kTier != FMGI_TIER_GRID && kTier != kTierRooms) soup = stage_soup(p, smem);

    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    // photon state
    bool alive = false, is_new = false, mirror = false;
    float px = 0, py = 0, pz = 0, dx = 0, dy = 0, dz = 1;
    float cr = 0, cg = 0, cb = 0, roulette = 0;
    int depth = 0, emitter = 0, hit_id = 0, leaf = 0;            // leaf: the photon's box (room tier)
    bool walking = false;                                        // room tier: the lane's ray is between boxes
    float ix = 0, iy = 0, iz = 0;                                // room tier: rooms_inv of the direction
    int boxes = 0;                                               // room tier: boxes the current ray has crossed
    unsigned long long photon = 0;
    // the warp's chunk (warp uniform): photon indices w_base + [w_pos, w_cnt) of emitter w_emitter
    unsigned long long w_base = 0;
    int w_pos = 0, w_cnt = 0, w_emitter = 0;
    bool exhausted = false;
    unsigned n_photons = 0, n_rays = 0, n_deposits = 0, n_mirror = 0, n_tests = 0;

    for (;;) {
        // ---- A. refill dead lanes from the warp's chunk of the photon index space ---------------
        // Fast path: the chunk still holds a photon for every dead lane - one ballot, lane l takes the
        // popc(dead lanes below l)-th of them.  Otherwise (chunk boundary, about once per chunk) the loop
        // below hands out what is left, claims the next chunk and goes on.
        is_new = false;
        unsigned dead = __ballot_sync(kFullMask, !alive);
        if (dead != 0u && !exhausted) {
            if (w_cnt - w_pos >= __popc(dead)) {
                if (!alive) {
                    photon = w_base + (unsigned)(w_pos + __popc(dead & lt_mask));
                    emitter = w_emitter;
                    alive = true; is_new = true; mirror = false; depth = 0;
                    n_photons++;
                }
                w_pos += __popc(dead);
                dead = 0u;
            } else {
                while (dead) {
                    if (w_pos == w_cnt) {
                        unsigned long long k = 0;
                        if (lane == 0) k = atomicAdd(p.work_counter, 1ull);
                        k = __shfl_sync(kFullMask, k, 0);
                        if (k >= p.total_jobs) { exhausted = true; break; }
                        w_emitter = find_emitter(p.job_begin, p.num_emitters, k);
                        const unsigned long long first = (k - __ldg(p.job_begin + w_emitter)) * (unsigned)p.chunk;
                        const unsigned long long left = __ldg(p.photon_count + w_emitter) - first;
                        w_base = __ldg(p.photon_first + w_emitter) + first;
                        w_cnt = left < (unsigned long long)p.chunk ? (int)left : p.chunk;
                        w_pos = 0;
                    }
                    const int avail = w_cnt - w_pos;
                    const int rank = __popc(dead & lt_mask);
                    if (!alive && rank < avail) {
                        photon = w_base + (unsigned)(w_pos + rank);
                        emitter = w_emitter;
                        alive = true; is_new = true; mirror = false; depth = 0;
                        n_photons++;
                    }
                    w_pos += min(__popc(dead), avail);
                    dead = __ballot_sync(kFullMask, !alive);
                }
            }
        }
        if (dead == kFullMask) break;

        bool dep = false;
        int idx = 0;
        if (alive && !(kTier == kTierRooms && walking)) {
            // ---- P. one Philox2x32 block per event: emission direction (event 0) or the bounce just done ----
            const uint32_t idw = philox_event_word((uint32_t)(photon >> 32), (uint32_t)emitter, 0u);
            const Philox2 w = philox2x32_10((uint32_t)photon, idw | ((uint32_t)depth << 28), p.philox_keys);
            // ---- S. new direction: emission and re-emission share the sampler ------------------------
            if (!FMGI_CHECK(p, is_new ? (unsigned)emitter < (unsigned)p.num_emitters : (unsigned)hit_id < p.num_walls, 7)) {
                emitter = 0; hit_id = 0;
            }
            const float4 *frame = is_new ? p.emitters + 6 * emitter : p.shade + 6 * hit_id;
            const float4 fn = ldg4(frame + 3);
            float4 e0 = make_float4(px, py, pz, 0.0f);
            if (is_new) {
                e0 = ldg4(frame);
                const bool sky = __float_as_int(e0.w) != 0;
                cr = sky ? 18.0f : 16.0f; cg = cr; cb = 18.0f;   // photonmap.c:169-171
            }
            if (mirror) {                                       // photonmap.c:230
                const float k2 = 2.0f * (fn.x * dx + fn.y * dy + fn.z * dz);
                dx = fmaf(-k2, fn.x, dx); dy = fmaf(-k2, fn.y, dy); dz = fmaf(-k2, fn.z, dz);
            } else {                                            // photonmap.c:179-181, :233
                float4 fu, fv;
                ldg256(frame + 4, fu, fv);
                sample_hemisphere(u24(w.w0), u24(w.w1), is_new && __float_as_int(e0.w) != 0, fn, fu, fv, dx, dy, dz);
            }
            roulette = r16(w.w0, w.w1);
            px = __fadd_rn(e0.x, __fmul_rn(dx, 1E-5f));         // photonmap.c:183, :254
            py = __fadd_rn(e0.y, __fmul_rn(dy, 1E-5f));
            pz = __fadd_rn(e0.z, __fmul_rn(dz, 1E-5f));
            if (is_new) {                                       // photonmap.c:175-176, :184-185
                const Philox2 wp = philox2x32_10((uint32_t)photon, idw | (kEventEmitPosition << 28), p.philox_keys);
                const float4 e1 = ldg4(frame + 1), e2 = ldg4(frame + 2);
                const float sx = u24(wp.w0), sy = u24(wp.w1);
                px = __fadd_rn(__fadd_rn(px, __fmul_rn(e1.x, sx)), __fmul_rn(e2.x, sy));
                py = __fadd_rn(__fadd_rn(py, __fmul_rn(e1.y, sx)), __fmul_rn(e2.y, sy));
                pz = __fadd_rn(__fadd_rn(pz, __fmul_rn(e1.z, sx)), __fmul_rn(e2.z, sy));
            }

            if (kTier == kTierRooms) {
                if (is_new) leaf = rooms_start(p, emitter, px, py, pz, dx, dy, dz);
                ix = rooms_inv(dx); iy = rooms_inv(dy); iz = rooms_inv(dz);
                boxes = 0;
            }
            n_rays++;
        }
        if (alive) {
            // ---- C. closest hit (photonmap.c:198 / photonmap.cl:194-206) ---------------------------
            float t;
            if (kTier == FMGI_TIER_SOUP) hit_id = closest_hit_soup(soup, px, py, pz, dx, dy, dz, t);
            else if (kTier == kTierSoupPlanes) hit_id = closest_hit_soup_planes<kCount>(soup, p, px, py, pz, dx, dy, dz, t, n_tests);
            else if (kTier == kTierRooms) {
                hit_id = rooms_walk<kCount, kRoomSteps>(p, leaf, px, py, pz, dx, dy, dz, ix, iy, iz, t, n_tests);
                boxes += kRoomSteps;
                if (hit_id == kRoomWalking && boxes >= kRoomMaxSteps) hit_id = -1;
                walking = hit_id == kRoomWalking;
            }
            else hit_id = closest_hit_grid<kCount>(p, px, py, pz, dx, dy, dz, t, n_tests);

            // ---- D. bounce: texel, roulette, attenuation (photonmap.c:200-247) ------------------------
            if (kTier == kTierRooms && walking) {
                // between boxes: nothing to shade yet
            } else if (hit_id < 0 || !FMGI_CHECK(p, (unsigned)hit_id < p.num_walls, 8)) {
                alive = false;                                   // photonmap.c:200-201: photon leaves the flat
            } else {
                px = __fadd_rn(px, __fmul_rn(dx, t));            // photonmap.c:208
                py = __fadd_rn(py, __fmul_rn(dy, t));
                pz = __fadd_rn(pz, __fmul_rn(dz, t));
                const float4 *sh = p.shade + 6 * hit_id;
                float4 q0, q1, q2, q3;
                ldg256(sh, q0, q1);
                ldg256(sh + 2, q2, q3);
                idx = tile_index(q0, q1, q2, __float_as_int(q3.w), px, py, pz);   // photonmap.c:210-211
                // floor is slightly reflective: photonmap.c:228 (0.0005 is a double there; for a float
                // z the comparison is equivalent to z < 0.0005f)
                mirror = pz < 0.0005f && roulette < 0.75f;
                if (!mirror) {
                    if (pz < 1E-5f) { cg *= 0.85f; cb *= 0.7f; } // photonmap.c:236-246
                    cr *= 0.9f; cg *= 0.9f; cb *= 0.9f;          // photonmap.c:247
                } else {
                    n_mirror++;
                }
                dep = FMGI_CHECK(p, (unsigned)idx < p.num_texels, 9);
                if (kProbe) p.path_out[(photon - __ldg(p.photon_first + emitter)) * p.max_depth + depth] = idx;
                depth++;
                n_deposits++;
                if (depth == p.max_depth) alive = false;         // photonmap.c:187
            }
        }
        // photonmap.c:251 — texels[idx] += colour, after attenuation
        if (!kProbe) deposit<kDeposit>(p.atlas, idx, cr, cg, cb, dep);
    }

    // ---- counters: warp reduce, one atomic per warp and counter --------------------------------------
    unsigned long long c0 = n_photons, c1 = n_rays, c2 = n_deposits, c3 = n_mirror, c5 = n_tests;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(kFullMask, c0, o);
        c1 += __shfl_xor_sync(kFullMask, c1, o);
        c2 += __shfl_xor_sync(kFullMask, c2, o);
        c3 += __shfl_xor_sync(kFullMask, c3, o);
        if (kTier != FMGI_TIER_SOUP) c5 += __shfl_xor_sync(kFullMask, c5, o);
    }
    if (lane == 0) {
        atomicAdd(p.counters + 0, c0);
        atomicAdd(p.counters + 1, c1);
        atomicAdd(p.counters + 2, c2);
        atomicAdd(p.counters + 3, c3);
        if (kTier != FMGI_TIER_SOUP) atomicAdd(p.counters + 5, c5);
    }
}

// ---- probe kernels: the same device functions, one item per thread --------------------------------------

template <int kTier>
__global__ void k_probe_closest_hit(const TraceParams p, const float *__restrict__ origins,
                                    const float *__restrict__ dirs, int num_rays, int32_t *hit_index, float *hit_dist)
{
    extern __shared__ float4 smem[];
    SoupTables soup;
    if (kTier != FMGI_TIER_GRID && kTier != kTierRooms) soup = stage_soup(p, smem);
    unsigned tests = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < num_rays; r += gridDim.x * blockDim.x) {
        float t;
        const float ox = origins[3 * r], oy = origins[3 * r + 1], oz = origins[3 * r + 2];
        const float dx = dirs[3 * r], dy = dirs[3 * r + 1], dz = dirs[3 * r + 2];
        const int id = kTier == kTierRooms ? closest_hit_rooms_from_anywhere(p, ox, oy, oz, dx, dy, dz, t)
                     : kTier == FMGI_TIER_SOUP ? closest_hit_soup(soup, ox, oy, oz, dx, dy, dz, t)
                     : kTier == kTierSoupPlanes ? closest_hit_soup_planes<false>(soup, p, ox, oy, oz, dx, dy, dz, t, tests)
                                                : closest_hit_grid<false>(p, ox, oy, oz, dx, dy, dz, t, tests);
        hit_index[r] = id;
        hit_dist[r] = t;
    }
}

__global__ void k_probe_tile_ids(const float4 *__restrict__ shade, const int32_t *__restrict__ rect_index,
                                 const float *__restrict__ points, int n, int32_t *tile_ids)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 *sh = shade + 6 * rect_index[i];
        const float4 q0 = sh[0], q1 = sh[1], q2 = sh[2], q3 = sh[3];
        tile_ids[i] = tile_index(q0, q1, q2, __float_as_int(q3.w), points[3 * i], points[3 * i + 1], points[3 * i + 2])
                      - __float_as_int(q0.w);
    }
}

__global__ void k_probe_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out)
{
    const Philox4 w = philox4x32_10(c0, c1, c2, c3, k0, k1);
    out[0] = w.w0; out[1] = w.w1; out[2] = w.w2; out[3] = w.w3;
}

__global__ void k_probe_philox2(uint32_t c0, uint32_t c1, uint32_t key, uint32_t *out)
{
    uint32_t keys[10];
    for (int r = 0; r < 10; r++) keys[r] = key + (uint32_t)r * kPhiloxW;
    const Philox2 w = philox2x32_10(c0, c1, keys);
    out[0] = w.w0; out[1] = w.w1;
}

// Directions as the emission event draws them: photon i of emitter 0, event 0.
__global__ void k_probe_sample_dirs(float4 n, float4 u, float4 v, int sky, uint32_t seed, int count, float *out)
{
    uint32_t keys[10];
    for (int r = 0; r < 10; r++) keys[r] = seed + (uint32_t)r * kPhiloxW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const Philox2 w = philox2x32_10((uint32_t)i, 0u, keys);
        float dx, dy, dz;
        sample_hemisphere(u24(w.w0), u24(w.w1), sky != 0, n, u, v, dx, dy, dz);
        out[3 * i] = dx; out[3 * i + 1] = dy; out[3 * i + 2] = dz;
    }
}

// Deposit roofline probe: the trace kernel's deposit instruction - one RED.E.ADD.F32x4 per
// bounce - at uniform-random texels of an atlas-sized footprint, with nothing else in the loop but a 32-bit
// mix that picks the texel.  Its rate is the "A_peak" the bench line compares the bake's deposit rate with.
__global__ void __launch_bounds__(256) k_probe_red_peak(float4 *__restrict__ atlas, uint32_t num_texels,
                                                        unsigned long long num_deposits, uint32_t seed)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < num_deposits; i += stride) {
        uint32_t h = (uint32_t)i * 0x9E3779B1u + seed;           // lowbias32-style mix: distinct per deposit
        h ^= h >> 16; h *= 0x21F0AAADu; h ^= h >> 15; h *= 0x735A2D97u; h ^= h >> 15;
        const uint32_t idx = (uint32_t)(((unsigned long long)h * num_texels) >> 32);
        atomicAdd(atlas + idx, make_float4(16.2f, 16.2f, 16.2f, 0.0f));
    }
}

// ---- ambient occlusion ------------------------------------------------------------------
//
// performAmbientOcclusionNative (photonmap.c:436-491) on the closest-hit code of the photon tracer: one
// thread per base-level texel, num_dirs rays from the texel centre (getTileCenter, rectangle.c:140-153)
// in the wall's local frame (createBase / transformToOrthoNormalBase, vector3_cl.c:152, photonmap.c:31),
// d = sum(dist * fac) / (1.5 * sum(fac)) with dist = 10 for a miss; the texel is overwritten with (d, d, d, 0).
struct AoWall {
    int32_t base;        // atlas index of the wall's base level
    int32_t first;       // index of its first texel in the flat texel enumeration
    int32_t wall;        // wall index (shading record)
    int32_t pad;
};

template <int kTier>
__global__ void __launch_bounds__(kTraceThreads) k_ambient_occlusion(const TraceParams p, const AoWall *__restrict__ walls,
                                                                     int num_walls, long long num_pixels,
                                                                     const float4 *__restrict__ dirs, int num_dirs)
{
    extern __shared__ float4 smem[];
    SoupTables soup;
    if (kTier != FMGI_TIER_GRID && kTier != kTierRooms) soup = stage_soup(p, smem);
    unsigned tests = 0;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < num_pixels;
         g += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = num_walls - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)walls[mid].first <= g) lo = mid; else hi = mid - 1;
        }
        const AoWall w = walls[lo];
        const int j = (int)(g - w.first);
        const float4 *sh = p.shade + 6 * w.wall;
        const float4 q0 = ldg4(sh), q1 = ldg4(sh + 1), q2 = ldg4(sh + 2), fn = ldg4(sh + 3), fu = ldg4(sh + 4), fv = ldg4(sh + 5);
        const int tw = __float_as_int(fn.w) & 0xffff, th = __float_as_int(fn.w) >> 16;
        // getTileCenter: pos + (width / tilesW) * (tx + 0.5) + (height / tilesH) * (ty + 0.5), with
        // width = wn * |width| recovered from the shading record
        const float rw = __fdiv_rn(1.0f, (float)tw), rh = __fdiv_rn(1.0f, (float)th);
        const float fx = (float)((double)(j % tw) + 0.5), fy = (float)((double)(j / tw) + 0.5);
        const float wx = __fmul_rn(p.ao_width[3 * w.wall], rw), wy = __fmul_rn(p.ao_width[3 * w.wall + 1], rw),
                    wz = __fmul_rn(p.ao_width[3 * w.wall + 2], rw);
        const float hx = __fmul_rn(p.ao_height[3 * w.wall], rh), hy = __fmul_rn(p.ao_height[3 * w.wall + 1], rh),
                    hz = __fmul_rn(p.ao_height[3 * w.wall + 2], rh);
        const float cx = __fadd_rn(__fadd_rn(q0.x, __fmul_rn(wx, fx)), __fmul_rn(hx, fy));
        const float cy = __fadd_rn(__fadd_rn(q0.y, __fmul_rn(wy, fx)), __fmul_rn(hy, fy));
        const float cz = __fadd_rn(__fadd_rn(q0.z, __fmul_rn(wz, fx)), __fmul_rn(hz, fy));
        (void)q1; (void)q2; (void)th;
        float dist_sum = 0.0f, fac_sum = 0.0f;
        // room tier: all rays of a texel start in the box in front of its wall - one tree descent per texel
        int box0 = 0;
        if (kTier == kTierRooms)
            box0 = rooms_locate(p, fmaf(fn.x, 1E-5f, cx), fmaf(fn.y, 1E-5f, cy), fmaf(fn.z, 1E-5f, cz), fn.x, fn.y, fn.z);
        for (int k = 0; k < num_dirs; k++) {
            const float4 d = __ldg(dirs + k);
            // photonmap.c:41-45: in.x * b0 + in.y * b1 + in.z * b2, (b0, b1, b2) = (U, V, n)
            const float dx = __fadd_rn(__fadd_rn(__fmul_rn(d.x, fu.x), __fmul_rn(d.y, fv.x)), __fmul_rn(d.z, fn.x));
            const float dy = __fadd_rn(__fadd_rn(__fmul_rn(d.x, fu.y), __fmul_rn(d.y, fv.y)), __fmul_rn(d.z, fn.y));
            const float dz = __fadd_rn(__fadd_rn(__fmul_rn(d.x, fu.z), __fmul_rn(d.y, fv.z)), __fmul_rn(d.z, fn.z));
            const float ox = __fadd_rn(cx, __fmul_rn(dx, 1E-5f)), oy = __fadd_rn(cy, __fmul_rn(dy, 1E-5f)),
                        oz = __fadd_rn(cz, __fmul_rn(dz, 1E-5f));                         // photonmap.c:457
            float t;
            int box = box0;
            const int id = kTier == kTierRooms ? closest_hit_rooms(p, box, ox, oy, oz, dx, dy, dz, t)
                         : kTier == FMGI_TIER_SOUP ? closest_hit_soup(soup, ox, oy, oz, dx, dy, dz, t)
                         : kTier == kTierSoupPlanes ? closest_hit_soup_planes<false>(soup, p, ox, oy, oz, dx, dy, dz, t, tests)
                                                    : closest_hit_grid<false>(p, ox, oy, oz, dx, dy, dz, t, tests);
            if (id < 0) t = 10.0f;                                                       // photonmap.c:462-466
            dist_sum = __fadd_rn(dist_sum, __fmul_rn(t, d.z));
            fac_sum = __fadd_rn(fac_sum, d.z);
        }
        const float v = (float)((double)dist_sum / ((double)fac_sum * 1.5));              // photonmap.c:473
        if (FMGI_CHECK(p, (unsigned)(w.base + j) < p.num_texels, 10)) p.atlas[w.base + j] = make_float4(v, v, v, 0.0f);
    }
}

// atlas += scratch (colour lanes only): folds one fp32 accumulation pass into the caller's atlas.
__global__ void k_accumulate(float4 *__restrict__ atlas, const float4 *__restrict__ scratch, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = atlas[i];
        const float4 b = scratch[i];
        a.x += b.x; a.y += b.y; a.z += b.z;
        atlas[i] = a;
    }
}

// Multi-GPU fold, reduce-scatter shaped: this GPU owns one slice of the atlas and sums that slice of every GPU's
// deposits - its own (`own`, in/out) and the peers', read straight over NVLink peer mappings - on top of the
// caller's values of the slice (`init`, all four lanes: lane 3 and the mip slots pass through untouched).
constexpr int kMaxFoldPeers = 31;
struct PeerList { const float4 *p[kMaxFoldPeers]; };

__global__ void k_fold_slice(float4 *__restrict__ own, const float4 *__restrict__ init, const PeerList peers, int num_peers,
                             size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 r = init[i];
        const float4 a = own[i];
        r.x += a.x; r.y += a.y; r.z += a.z;
        for (int g = 0; g < num_peers; g++) {
            const float4 b = peers.p[g][i];
            r.x += b.x; r.y += b.y; r.z += b.z;
        }
        own[i] = r;
    }
}

}  // namespace fmgi
