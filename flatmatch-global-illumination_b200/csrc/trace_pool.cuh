// Pooled trace kernel for the grid tier: the same photon loop as k_trace (trace_core.cuh), re-cut so that the
// closest-hit walk - 46 % of k_trace's instructions at 12 of 32 lanes, because a warp waits for its longest of 32
// walks (mean 6 steps, warp maximum 16: profiles/r2_walk_histogram.md) - runs over a POOL of K * 32 rays per warp:
//
//   phase SG (full lanes, K batches of 32 slots): shade the hit the last walk found for the slot's photon
//     (z planes, texel index, roulette, attenuation, deposit), refill dead slots from the warp's photon chunk, then
//     draw the photon's next ray (Philox, hemisphere sample) and store it walk-ready in shared memory;
//   phase W: every lane walks one ray; a lane whose walk is over stores the result and takes the next ray of the
//     warp's pool, so the 35-instruction walk loop keeps ~21-25 lanes busy instead of 12.  The loop is the predicated
//     PTX of GridWalk::walk with a warp-uniform exit: it runs while at least (32 - kSwitchLanes) lanes are still
//     walking (or, once the pool is empty, until the last walk ends); a finished lane re-executes its last iteration,
//     which is idempotent (its pending record fails t < best, its DDA does not step, its load is predicated off).
//
// Everything is warp-private: no __syncthreads, photon and ray state live in a per-warp slice of shared memory
// (80 bytes per slot).  The random streams are keyed by (seed, emitter, photon, event), so the sample set - and
// with it counters and atlas, up to the order of the float atomics - is the one k_trace produces.
// Scenes whose walk lists hold misc records, the soup tiers, probes and the counting variant stay on k_trace.
#pragma once
#include "trace_core.cuh"

namespace fmgi {

constexpr int kPoolThreads = 128;           // 4 warps per CTA: 40 KB of shared memory at K = 4
constexpr int kPoolSlotBytes = 80;          // five float4 per slot
constexpr int kSwitchLanes = 8;             // idle lanes that trigger a switch while the pool has rays

// the slot's packed photon id word is the Philox counter word (philox.cuh): same limits
constexpr unsigned long long kPoolMaxPhotonIndex = kPhiloxMaxPhotons;
constexpr int kPoolMaxEmitters = kPhiloxMaxEmitters;
constexpr int kPoolMaxDepth = kPhiloxMaxDepth;

// Walk set-up of one ray: what GridWalk::walk computes before its loop (same expressions).
struct WalkStart {
    float tmx, tmy, t_exit, ax, ay;
    int ci;
};

__device__ __forceinline__ WalkStart walk_setup(const TraceParams &p, float ox, float oy, float oz, float dx, float dy,
                                                float dz)
{
    const GridDesc &g = p.grid;
    const float inf = __int_as_float(0x7f800000);
    const bool xp = dx > 0.0f, yp = dy > 0.0f;
    const bool x0 = dx == 0.0f, y0 = dy == 0.0f;
    const float ix = rcp_fast(dx), iy = rcp_fast(dy), iz = rcp_fast(dz);
    int cx = __float2int_rd(fmaf(ox, g.inv_cell, g.bx)), cy = __float2int_rd(fmaf(oy, g.inv_cell, g.by));
    cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1);
    float tmx = (fmaf((float)(cx + (xp ? 1 : 0)), g.cell, g.x0) - ox) * ix;
    float tmy = (fmaf((float)(cy + (yp ? 1 : 0)), g.cell, g.y0) - oy) * iy;
    const float ex = ((xp ? g.exit_hi_x : g.exit_lo_x) - ox) * ix, ey = ((yp ? g.exit_hi_y : g.exit_lo_y) - oy) * iy;
    const float ez = ((dz < 0.0f ? g.wall_z_lo : g.wall_z_hi) - oz) * iz * 1.0001f;
    WalkStart w;
    w.tmx = x0 ? inf : fmaxf(tmx, 0.0f); w.tmy = y0 ? inf : fmaxf(tmy, 0.0f);
    w.t_exit = fmaxf(fminf(fminf(x0 ? inf : ex, y0 ? inf : ey), dz == 0.0f ? inf : ez), 0.0f);
    w.ci = g.walk_base + ((xp ? 1 : 0) + (yp ? 2 : 0)) * g.ncell + cy * g.nx + cx;
    const float nanv = __int_as_float(0x7fc00000);
    w.ax = x0 ? nanv : ix; w.ay = y0 ? nanv : iy;
    return w;
}

// Phase W: walks the warp's P walk-ready rays.  wa = {o, tmx}, wb = {d, tmy}, wc = {t_exit, ci, ax, ay}; the result
// (ray parameter of the wall hit or +inf, index of its record in T or -1) replaces wa.w / wb.w.
__device__ __forceinline__ void walk_pool(const TraceParams &p, float4 *wa, float4 *wb, const float4 *wc, int P, int lane)
{
    const GridDesc &g = p.grid;
    const float inf = __int_as_float(0x7f800000);
    const unsigned lt_mask = (1u << lane) - 1u;
    int next = 0;                                     // warp uniform: first slot nobody has taken yet
    int slot = 0;
    bool have = false;
    // a lane without a ray holds the null walk: the dummy head of a margin cell, every bound 0 - one idempotent step
    float best = 0.0f, tmx = 0.0f, tmy = 0.0f;
    int win = -1, r = 0, rend = 0, ci = g.walk_base, cur = g.walk_base;
    float4 q0 = make_float4(0.0f, -1.0f, 0.0f, -1.0f);
    float qc = __int_as_float(0x7fc00000);
    unsigned qtag = 0, tests = 0;
    float ax = qc, ay = qc, bx = 0.0f, by = 0.0f, ox = 0.0f, oy = 0.0f, oz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 0.0f;
    int sx = 0, sy = 0, more = 0, misc = 0;
    for (;;) {
        if (!more && have) {                          // this lane's walk is over: hand the result to phase SG
            wa[slot].w = win < 0 ? inf : best;
            wb[slot].w = __int_as_float(win);
            have = false;
        }
        const unsigned idle = __ballot_sync(kFullMask, !more);
        const int avail = P - next;
        if (idle && avail > 0) {
            const int rank = __popc(idle & lt_mask);
            if (!more && rank < avail) {
                slot = next + rank;
                const float4 a = wa[slot], b = wb[slot];
                float4 c = wc[slot];
                if (!FMGI_CHECK(p, (unsigned)__float_as_int(c.y) < p.grid_records, 11)) c.y = __int_as_float(g.walk_base);
                ox = a.x; oy = a.y; oz = a.z; tmx = a.w;
                dx = b.x; dy = b.y; dz = b.z; tmy = b.w;
                best = c.x; ci = __float_as_int(c.y); ax = c.z; ay = c.w;
                bx = -ox * ax; by = -oy * ay;
                sx = dx > 0.0f ? 1 : -1; sy = dy > 0.0f ? g.nx : -g.nx;
                float4 h1;
                ldg256(p.grid_table + 2 * ci, q0, h1);
                qc = h1.x; qtag = __float_as_uint(h1.y);
                r = __float_as_int(h1.z); rend = __float_as_int(h1.w);
                cur = ci; win = -1; have = true; more = 1;
            }
            next += min(__popc(idle), avail);
        }
        if (__ballot_sync(kFullMask, more) == 0u) break;          // pool empty, every walk over
        // keep walking while this many lanes still have steps to do
        const int keep = next < P ? 32 - kSwitchLanes + 1 : 1;
        // GridWalk::walk's loop (trace_kernels.cuh, balanced form) with the warp-uniform exit described above
        asm volatile(
            "{\n\t"
            ".reg .pred ky, ok, adv, cont, stepx, go, gx, gy, more, again;\n\t"
            ".reg .f32 ak, bk, dh, oh, t, pi, pj, tn, sa;\n\t"
            ".reg .b32 tb, bb, vb;\n\t"
            ".reg .b64 a;\n\t"
            "PWALK:\n\t"
            "setp.lt.s32 ky, %13, 0;\n\t"
            "selp.f32 ak, %20, %19, ky;\n\t"
            "selp.f32 bk, %22, %21, ky;\n\t"
            "add.rn.f32 dh, %27, 0f80000000;\n\t"
            "@ky add.rn.f32 dh, %26, 0f80000000;\n\t"
            "add.rn.f32 oh, %24, 0f80000000;\n\t"
            "@ky add.rn.f32 oh, %23, 0f80000000;\n\t"
            "fma.rn.f32 t, %12, ak, bk;\n\t"
            "fma.rn.f32 pi, t, dh, oh;\n\t"
            "fma.rn.f32 pj, t, %28, %25;\n\t"
            "sub.rn.f32 pi, pi, %8;\n\t"
            "sub.rn.f32 pj, pj, %10;\n\t"
            "abs.f32 pi, pi;\n\t"
            "abs.f32 pj, pj;\n\t"
            "mov.b32 tb, t;\n\t"
            "mov.b32 bb, %0;\n\t"
            "setp.lt.u32 ok, tb, bb;\n\t"
            "setp.le.and.f32 ok, pi, %9, ok;\n\t"
            "setp.le.and.f32 ok, pj, %11, ok;\n\t"
            "@ok add.rn.f32 %0, t, 0f80000000;\n\t"
            "@ok mad.lo.s32 %1, %14, %17, 0;\n\t"
            "setp.ge.s32 adv, %2, %3;\n\t"
            "min.f32 tn, %5, %6;\n\t"
            "setp.lt.f32 cont, tn, %0;\n\t"
            "setp.lt.f32 stepx, %5, %6;\n\t"
            "and.pred go, adv, cont;\n\t"
            "and.pred gx, go, stepx;\n\t"
            "and.pred gy, go, !stepx;\n\t"
            "or.pred more, cont, !adv;\n\t"
            "@gx mad.lo.s32 %4, %31, %17, %4;\n\t"
            "@gy mad.lo.s32 %4, %32, %17, %4;\n\t"
            "abs.f32 sa, %19;\n\t"
            "@gx fma.rn.f32 %5, sa, %29, %5;\n\t"
            "abs.f32 sa, %20;\n\t"
            "@gy fma.rn.f32 %6, sa, %29, %6;\n\t"
            "selp.b32 %14, %4, %2, go;\n\t"
            "mul.wide.s32 a, %14, 32;\n\t"
            "add.s64 a, a, %18;\n\t"
            "@more ld.global.nc.v8.b32 {%8, %9, %10, %11, %12, %13, %2, %3}, [a];\n\t"
            "selp.u32 %15, 1, 0, more;\n\t"
            "vote.sync.ballot.b32 vb, more, 0xffffffff;\n\t"
            "popc.b32 vb, vb;\n\t"
            "setp.ge.s32 again, vb, %33;\n\t"
            "@again bra PWALK;\n\t"
            "}"
            : "+f"(best), "+r"(win), "+r"(r), "+r"(rend), "+r"(ci), "+f"(tmx), "+f"(tmy), "+r"(tests), "+f"(q0.x),
              "+f"(q0.y), "+f"(q0.z), "+f"(q0.w), "+f"(qc), "+r"(qtag), "+r"(cur), "=r"(more), "=r"(misc)
            : "r"(p.one), "l"(p.grid_table), "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(ox), "f"(oy), "f"(oz), "f"(dx),
              "f"(dy), "f"(dz), "f"(g.cell), "f"(0.0f), "r"(sx), "r"(sy), "r"(keep));
    }
}

// K: rays per lane in the pool (pool = 32 * K slots per warp).
template <int K, int kDeposit, int kMinBlocks>
__global__ void __launch_bounds__(kPoolThreads, kMinBlocks) k_trace_pool(const TraceParams p)
{
    extern __shared__ float4 smem[];
    constexpr int P = 32 * K;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    // the warp's slice: ray records wa = {o, tmx | hit distance}, wb = {d, tmy | hit record}, wc = {t_exit, ci, ax, ay};
    // photon state sa = {photon lo, photon hi | emitter << 8 | bounce << 28, roulette, red}, sb = {green, blue, alive, -}
    float4 *wa = smem + (threadIdx.x >> 5) * (5 * P), *wb = wa + P, *wc = wb + P, *sa = wc + P, *sb = sa + P;
    const float inf = __int_as_float(0x7f800000);
    const float nanv = __int_as_float(0x7fc00000);

#pragma unroll
    for (int j = 0; j < K; j++) sb[j * 32 + lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);     // every slot starts dead

    // the warp's chunk (warp uniform): photon indices w_base + [w_pos, w_cnt) of emitter w_emitter
    unsigned long long w_base = 0;
    int w_pos = 0, w_cnt = 0, w_emitter = 0;
    bool exhausted = false;
    unsigned n_photons = 0, n_rays = 0, n_deposits = 0, n_mirror = 0;

    for (;;) {
        // ---- phase SG: shade the last hit of every slot, refill, draw the next ray ----------------------------
        unsigned any_alive = 0;
#pragma unroll 1
        for (int j = 0; j < K; j++) {
            const int s = j * 32 + lane;
            const float4 st_a = sa[s], st_b = sb[s];
            uint32_t photon_lo = __float_as_uint(st_a.x), idw = __float_as_uint(st_a.y);
            float roulette = st_a.z, cr = st_a.w, cg = st_b.x, cb = st_b.y;
            bool alive = st_b.z != 0.0f, mirror = false, is_new = false;
            int depth = (int)(idw >> 28), emitter = (int)((idw >> 8) & 0xfffffu), hit_id = 0;
            uint32_t photon_hi = idw & 0xffu;
            float px = 0, py = 0, pz = 0, dx = 0, dy = 0, dz = 1;

            // ---- D. the hit of the ray this slot held (k_trace sections C-end and D) ------------------------
            bool dep = false;
            int idx = 0;
            if (alive) {
                const float4 ra = wa[s], rb = wb[s];
                px = ra.x; py = ra.y; pz = ra.z; dx = rb.x; dy = rb.y; dz = rb.z;
                GridWalk w;
                w.best = ra.w; w.win = __float_as_int(rb.w);
                unsigned tests = 0;
                w.template planes<false>(p, px, py, pz, dx, dy, dz, tests);
                float t;
                hit_id = w.finish(p, px, py, pz, dx, dy, dz, t);
                n_rays++;
                if (hit_id < 0 || !FMGI_CHECK(p, (unsigned)hit_id < p.num_walls, 8)) {
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
                    mirror = pz < 0.0005f && roulette < 0.75f;       // photonmap.c:228
                    if (!mirror) {
                        if (pz < 1E-5f) { cg *= 0.85f; cb *= 0.7f; } // photonmap.c:236-246
                        cr *= 0.9f; cg *= 0.9f; cb *= 0.9f;          // photonmap.c:247
                    } else {
                        n_mirror++;
                    }
                    dep = FMGI_CHECK(p, (unsigned)idx < p.num_texels, 9);
                    depth++;
                    n_deposits++;
                    if (depth == p.max_depth) alive = false;         // photonmap.c:187
                }
            }
            deposit<kDeposit>(p.atlas, idx, cr, cg, cb, dep);        // photonmap.c:251, after attenuation

            // ---- A. refill dead slots of this batch from the warp's chunk (as k_trace) ---------------------------
            if (!exhausted) {
                unsigned dead = __ballot_sync(kFullMask, !alive);
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
                        const unsigned long long photon = w_base + (unsigned)(w_pos + rank);
                        photon_lo = (uint32_t)photon; photon_hi = (uint32_t)(photon >> 32);
                        emitter = w_emitter;
                        alive = true; is_new = true; mirror = false; depth = 0;
                        n_photons++;
                    }
                    w_pos += min(__popc(dead), avail);
                    dead = __ballot_sync(kFullMask, !alive);
                }
            }
            any_alive |= __ballot_sync(kFullMask, alive);

            // ---- P, S. the photon's next ray (k_trace sections P and S), stored walk-ready --------------------------
            float4 na = make_float4(0.0f, 0.0f, 0.0f, 0.0f), nb = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
            float4 nc = make_float4(0.0f, __int_as_float(p.grid.walk_base), nanv, nanv);          // the null walk
            if (alive) {
                const uint32_t idw0 = philox_event_word(photon_hi, (uint32_t)emitter, 0u);
                const Philox2 w = philox2x32_10(photon_lo, idw0 | ((uint32_t)depth << 28), p.philox_keys);
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
                    const Philox2 wp = philox2x32_10(photon_lo, idw0 | (kEventEmitPosition << 28), p.philox_keys);
                    const float4 e1 = ldg4(frame + 1), e2 = ldg4(frame + 2);
                    const float ux = u24(wp.w0), uy = u24(wp.w1);
                    px = __fadd_rn(__fadd_rn(px, __fmul_rn(e1.x, ux)), __fmul_rn(e2.x, uy));
                    py = __fadd_rn(__fadd_rn(py, __fmul_rn(e1.y, ux)), __fmul_rn(e2.y, uy));
                    pz = __fadd_rn(__fadd_rn(pz, __fmul_rn(e1.z, ux)), __fmul_rn(e2.z, uy));
                }
                const WalkStart ws = walk_setup(p, px, py, pz, dx, dy, dz);
                na = make_float4(px, py, pz, ws.tmx);
                nb = make_float4(dx, dy, dz, ws.tmy);
                nc = make_float4(ws.t_exit, __int_as_float(ws.ci), ws.ax, ws.ay);
            }
            wa[s] = na; wb[s] = nb; wc[s] = nc;
            sa[s] = make_float4(__uint_as_float(photon_lo),
                                __uint_as_float(photon_hi | ((uint32_t)emitter << 8) | ((uint32_t)depth << 28)), roulette, cr);
            sb[s] = make_float4(cg, cb, alive ? 1.0f : 0.0f, 0.0f);
        }
        if (any_alive == 0u) break;
        __syncwarp();
        // ---- phase W: walk the pool ---------------------------------------------------------------------------
        walk_pool(p, wa, wb, wc, P, lane);
        __syncwarp();
    }
    (void)inf;

    // ---- counters: warp reduce, one atomic per warp and counter --------------------------------------
    unsigned long long c0 = n_photons, c1 = n_rays, c2 = n_deposits, c3 = n_mirror;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(kFullMask, c0, o);
        c1 += __shfl_xor_sync(kFullMask, c1, o);
        c2 += __shfl_xor_sync(kFullMask, c2, o);
        c3 += __shfl_xor_sync(kFullMask, c3, o);
    }
    if (lane == 0) {
        atomicAdd(p.counters + 0, c0);
        atomicAdd(p.counters + 1, c1);
        atomicAdd(p.counters + 2, c2);
        atomicAdd(p.counters + 3, c3);
    }
}

}  // namespace fmgi
