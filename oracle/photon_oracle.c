/* TEST INFRASTRUCTURE — CPU restatement of the reference's photon-mapping path.
 * See photon_oracle.h for the role and the parity pin.  Every function cites the reference
 * lines it restates.  Compiled with -ffp-contract=off: the reference is built for plain SSE
 * (no FMA), so every float expression below keeps the reference's operation order and
 * rounding points.  Plain C99, libm only.
 */
#include "photon_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small float3 helpers (vector3_cl.c:8-100).  All sums are left-associated like the reference.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld(const float *p) { return V(p[0], p[1], p[2]); }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float vsqlen(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline float vlen(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* vector3_cl.c:53-58: division is a multiply by the rounded reciprocal */
static inline v3 vdiv(v3 a, float b) { float rec = 1.0f / b; return V(a.x * rec, a.y * rec, a.z * rec); }
/* vector3_cl.c:95-100 */
static inline v3 vnormalized(v3 a) { float fac = 1.0f / vlen(a); return V(a.x * fac, a.y * fac, a.z * fac); }
/* vector3_cl.c:78-83 */
static inline v3 vcross(v3 a, v3 b)
{
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* vector3_cl.c:24-29 */
static inline v3 vadd3(v3 a, v3 b, v3 c) { return V(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z); }
/* vector3_cl.c:31-36 */
static inline v3 vadd4(v3 a, v3 b, v3 c, v3 d)
{
    return V(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z);
}

/* ------------------------------------------------------------------------------------------
 * rectangle.c:67-95 — ray/rectangle test with back-face culling and inclusive edges
 * ---------------------------------------------------------------------------------------- */
static float intersects_v(const orc_rect *rect, v3 src, v3 dir, float closest)
{
    v3 n = ld(rect->n), pos = ld(rect->pos), w = ld(rect->width), h = ld(rect->height);
    float denom = vdot(n, dir);
    if (denom >= 0)
        return -1;
    float fac = vdot(n, vsub(pos, src)) / denom;
    if (fac < 0)
        return -1;
    v3 ray = vmul(dir, fac);
    if (closest * closest < vsqlen(ray))
        return -1;
    v3 pdir = vsub(vadd(src, ray), pos);
    float wlen = vlen(w), hlen = vlen(h);
    float dx = vdot(vdiv(w, wlen), pdir);
    float dy = vdot(vdiv(h, hlen), pdir);
    if (dx < 0 || dy < 0 || dx > wlen || dy > hlen)
        return -1;
    return fac;
}

float orc_intersects(const orc_rect *rect, const float src[3], const float dir[3], float closest)
{
    return intersects_v(rect, ld(src), ld(dir), closest);
}

/* ------------------------------------------------------------------------------------------
 * rectangle.c:197-230 — point on a rectangle -> texel index inside its base-level tile grid
 * ---------------------------------------------------------------------------------------- */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static int tile_id_v(const orc_rect *rect, v3 p)
{
    v3 pdir = vsub(p, ld(rect->pos));
    v3 w = ld(rect->width), h = ld(rect->height);
    float hlen = vlen(w), vlen_ = vlen(h);
    float dx = vdot(vdiv(w, hlen), pdir);
    float dy = vdot(vdiv(h, vlen_), pdir);
    int nh = rect->lm[1], nv = rect->lm[2];
    int tx = clampi((int)(dx * nh / hlen), 0, nh - 1);
    int ty = clampi((int)(dy * nv / vlen_), 0, nv - 1);
    return ty * nh + tx;
}

int orc_tile_id(const orc_rect *rect, const float p[3]) { return tile_id_v(rect, ld(p)); }

/* photonmap.c:414-418 / 423-430 */
uint64_t orc_photon_budget(const orc_rect *e, int spa)
{
    float area = vlen(ld(e->width)) * vlen(ld(e->height));
    return (uint64_t)(spa * area);
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10.  Not in the reference (which uses libc rand(), photonmap.c:175): this is the
 * CUDA path's generator restated so that photon paths can be compared one by one.
 * ---------------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Philox2x32-10 (same paper; M = D256D193, W = 9E3779B9): the generator the CUDA tracer draws from. */
void orc_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2])
{
    uint32_t c0 = ctr[0], c1 = ctr[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD256D193u * c0;
        uint32_t n0 = (uint32_t)(p0 >> 32) ^ key ^ c1;
        c1 = (uint32_t)p0; c0 = n0;
        key += 0x9E3779B9u;
    }
    out[0] = c0; out[1] = c1;
}

/* ------------------------------------------------------------------------------------------
 * Random source.  LIBC: rand()/(double)RAND_MAX in call order (photonmap.c:175-176,228;
 * vector3_cl.c:107-108,131-132).
 * PHILOX (the CUDA path's stream, see DESIGN.md "Random stream"): one Philox2x32-10 block per event,
 *   key = seed, counter = {photon lo, photon hi (8 bits) | emitter << 8 | event << 28};
 *   u24(w) = (w >> 8) * 2^-24;  r16(a, b) = (((a & 255) << 8) | (b & 255)) * 2^-16;
 *   event 15 (emission position): dx = u24(w0), dy = u24(w1);
 *   event 0 (emission direction): xi1 = u24(w0), xi2 = u24(w1), roulette of bounce 1 = r16(w0, w1);
 *   event b (after bounce b): xi1 = u24(w0), xi2 = u24(w1), roulette of bounce b+1 = r16(w0, w1).
 * The roulette draw of a bounce is therefore known before the bounce happens, which lets the
 * kernel finish a photon's last deposit without another generator call.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int mode;
    uint32_t key[2];
    uint32_t photon_lo, photon_hi;
    uint32_t block[4];
    double roulette_next;
    uint64_t seq;              /* ORC_RNG_SEQ31 state */
} rng_t;

static void rng_event(rng_t *g, uint32_t event)
{
    if (g->mode == ORC_RNG_PHILOX) {
        uint32_t ctr[2] = {g->photon_lo, g->photon_hi | (g->key[1] << 8) | (event << 28)};   /* key[1] = emitter */
        orc_philox2x32_10(ctr, g->key[0], g->block);
        uint32_t a = g->block[0], b = g->block[1];
        if (event != 15)
            g->roulette_next = (double)((float)(((a & 255u) << 8) | (b & 255u)) * (1.0f / 65536.0f));
    }
}

static double rng_u01(rng_t *g, int slot)
{
    if (g->mode == ORC_RNG_LIBC)
        return rand() / (double)RAND_MAX;
    if (g->mode == ORC_RNG_SEQ31) {            /* splitmix64 (Steele, Lea, Flood 2014), top 31 bits */
        uint64_t z = (g->seq += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 33) / 2147483647.0;
    }
    return (double)((float)(g->block[slot] >> 8) * (1.0f / 16777216.0f));
}

/* ------------------------------------------------------------------------------------------
 * vector3_cl.c:102-149 — Malley disk sampling around ndir; `sky` folds u onto +U
 * (vector3_cl.c:115-116).  sqrt/cos/sin are evaluated in double and rounded to float, as the
 * reference's implicit promotions do.  slot0/slot1 are the PHILOX word positions of xi1, xi2.
 * ---------------------------------------------------------------------------------------- */
static v3 sample_hemisphere(rng_t *g, v3 ndir, int sky, int slot0)
{
    float r = sqrt(rng_u01(g, slot0));
    float phi = 2 * 3.141592f * rng_u01(g, slot0 + 1);
    float u = r * cos(phi);
    float v = r * sin(phi);
    float n = sqrt(1 - r * r);
    if (sky && u < 0)
        u = -u;
    v3 udir = V(0, 0, 1);
    if (fabs(vdot(udir, ndir)) >= 0.999999f)
        udir = V(0, 1, 0);
    v3 vdir = vnormalized(vcross(udir, ndir));
    udir = vnormalized(vcross(vdir, ndir));
    return vadd3(vmul(udir, u), vmul(vdir, v), vmul(ndir, n));
}

/* ------------------------------------------------------------------------------------------
 * BSP tree — photonmap.c:20-27 (node), :278-300 (split cost), :302-374 (subdivide),
 * :388-406 (build).  Nodes hold indices into the caller's wall table instead of Rectangle
 * copies; item ORDER inside each node follows the reference's swap-remove exactly, because the
 * order decides both the split choice on ties and the scan order during traversal.
 * ---------------------------------------------------------------------------------------- */
typedef struct bsp_node {
    int plane;                     /* wall index of the split plane (valid if it has children) */
    int *items;
    int num_items;
    struct bsp_node *left, *right;
} bsp_node;

/* rectangle.c:436-440 */
static float dist_to_plane(const orc_rect *plane, v3 p)
{
    return vdot(vsub(p, ld(plane->pos)), ld(plane->n));
}

/* rectangle.c:476-506: -1 / +1 if all four corners are on one side, else 0 */
static int side_of(const orc_rect *plane, const orc_rect *rect)
{
    v3 pos = ld(rect->pos), w = ld(rect->width), h = ld(rect->height);
    v3 c[4] = {pos, vadd(pos, w), vadd(pos, h), vadd3(pos, w, h)};
    int is_left = 0, is_right = 0;
    for (int i = 0; i < 4; i++) {
        double d = dist_to_plane(plane, c[i]);
        is_left |= (d < 0);
        is_right |= (d > 0);
    }
    if (is_left && !is_right) return -1;
    if (is_right && !is_left) return 1;
    return 0;
}

/* photonmap.c:278-300 */
static int split_cost(const orc_rect *walls, const bsp_node *node, int plane)
{
    int l = 0, r = 0, c = 0;
    for (int i = 0; i < node->num_items; i++) {
        int s = side_of(&walls[plane], &walls[node->items[i]]);
        l += (s < 0); r += (s > 0); c += (s == 0);
    }
    return (l > r ? l : r) + c;
}

/* photonmap.c:302-374 */
static void subdivide(const orc_rect *walls, bsp_node *node)
{
    if (node->num_items < 20)
        return;
    int lowest = node->num_items, split = 0;
    for (int i = 0; i < node->num_items; i++) {
        int cost = split_cost(walls, node, node->items[i]);
        if (cost < lowest) { lowest = cost; split = i; }
    }
    int *left = malloc(sizeof(int) * node->num_items), *right = malloc(sizeof(int) * node->num_items);
    int nl = 0, nr = 0;
    node->plane = node->items[split];
    for (int i = 0; i < node->num_items;) {
        int item = node->items[i];
        int s = side_of(&walls[node->plane], &walls[item]);
        if (s < 0) left[nl++] = item;
        if (s > 0) right[nr++] = item;
        if (s != 0) node->items[i] = node->items[--node->num_items];
        else i++;
    }
    if (nl) {
        node->left = calloc(1, sizeof(bsp_node));
        node->left->items = left; node->left->num_items = nl;
        subdivide(walls, node->left);
    } else free(left);
    if (nr) {
        node->right = calloc(1, sizeof(bsp_node));
        node->right->items = right; node->right->num_items = nr;
        subdivide(walls, node->right);
    } else free(right);
}

static bsp_node *bsp_build(const orc_rect *walls, int n)
{
    bsp_node *root = calloc(1, sizeof(bsp_node));
    root->items = malloc(sizeof(int) * (n > 0 ? n : 1));
    root->num_items = n;
    for (int i = 0; i < n; i++) root->items[i] = i;
    subdivide(walls, root);
    return root;
}

static void bsp_free(bsp_node *n)
{
    if (!n) return;
    bsp_free(n->left); bsp_free(n->right);
    free(n->items); free(n);
}

/* rectangle.c:115-129 */
static float plane_hit_dist(v3 src, v3 dir, v3 pn, v3 ppos)
{
    float denom = vdot(pn, dir);
    if (denom == 0) return -1;
    float fac = vdot(pn, vsub(ppos, src)) / denom;
    if (fac < 0) return -1;
    return fac;
}

/* photonmap.c:54-161 — near child first; far child only when the near side had no hit and the
 * ray faces the split plane, with the origin advanced onto the plane (dist_shift). */
static int bsp_closest(const orc_rect *walls, const bsp_node *node, v3 pos, v3 dir,
                       float *dist, float shift, int *target, uint64_t *tests)
{
    int has_hit = 0;
    for (int i = 0; i < node->num_items; i++) {
        float d = intersects_v(&walls[node->items[i]], pos, dir, *dist);
        (*tests)++;
        if (d == -1)
            continue;
        if (d + shift < *dist) { *target = node->items[i]; *dist = d + shift; has_hit = 1; }
    }
    if (!node->left && !node->right)
        return has_hit;

    const orc_rect *pl = &walls[node->plane];
    v3 sn = ld(pl->n);
    if (vdot(vsub(pos, ld(pl->pos)), sn) < 0)
        sn = vneg(sn);
    int faces_away = vdot(sn, dir) >= 0;
    float side = dist_to_plane(pl, pos);

    const bsp_node *near_ = side < 0 ? node->left : node->right;
    const bsp_node *far_ = side < 0 ? node->right : node->left;
    int child_hit = 0;
    if (near_)
        child_hit = bsp_closest(walls, near_, pos, dir, dist, shift, target, tests);
    if (!child_hit && far_ && !faces_away) {
        float pd = plane_hit_dist(pos, dir, ld(pl->n), ld(pl->pos));
        if (pd < 0) pd = 0;
        pos = vadd(pos, vmul(dir, pd));
        has_hit |= bsp_closest(walls, far_, pos, dir, dist, shift + pd, target, tests);
    }
    has_hit |= child_hit;
    return has_hit;
}

/* photonmap.cl:194-206 / debugRaytracer.cc:47-66 — linear scan, lowest index wins ties */
static void linear_closest(const orc_rect *walls, int n, v3 pos, v3 dir, float *dist, int *target,
                           uint64_t *tests)
{
    for (int i = 0; i < n; i++) {
        float d = intersects_v(&walls[i], pos, dir, *dist);
        (*tests)++;
        if (d < 0) continue;
        if (d < *dist) { *dist = d; *target = i; }
    }
}

typedef struct {
    const orc_rect *walls;
    int num_walls;
    int accel;
    bsp_node *root;
} scene_t;

static void closest(const scene_t *s, v3 pos, v3 dir, float *dist, int *target, uint64_t *tests)
{
    *dist = INFINITY; *target = -1;
    if (s->accel == ORC_ACCEL_BSP)
        bsp_closest(s->walls, s->root, pos, dir, dist, 0, target, tests);
    else
        linear_closest(s->walls, s->num_walls, pos, dir, dist, target, tests);
}

/* ------------------------------------------------------------------------------------------
 * photonmap.c:164-257 — one photon: emit, then up to max_depth x (closest hit, roulette,
 * attenuate, deposit, re-emit).  path (optional) records the atlas index of each deposit.
 * ---------------------------------------------------------------------------------------- */
static void trace_photon(const scene_t *s, const orc_rect *src, int is_window, rng_t *g,
                         int max_depth, float *texels, orc_stats *st, int32_t *path)
{
    v3 colour = is_window ? V(18, 18, 18) : V(16, 16, 18);           /* photonmap.c:169-171 */
    rng_event(g, 15);                                                /* PHILOX: emission position block */
    float dx = rng_u01(g, 0);                                        /* :175 */
    float dy = rng_u01(g, 1);                                        /* :176 */
    rng_event(g, 0);                                                 /* PHILOX: emission direction block */
    v3 dir = sample_hemisphere(g, ld(src->n), is_window, 0);         /* :179-181 */
    v3 pos = vadd4(ld(src->pos), vmul(dir, 1E-5f),                   /* :183-185 */
                   vmul(ld(src->width), dx), vmul(ld(src->height), dy));
    st->photons++;

    for (int depth = 0; depth < max_depth; depth++) {
        float dist; int hit;
        closest(s, pos, dir, &dist, &hit, &st->rect_tests);          /* :198 */
        st->rays++;
        if (dist == INFINITY)                                        /* :200 */
            return;
        const orc_rect *obj = &s->walls[hit];
        pos = vadd(pos, vmul(dir, dist));                            /* :208 */
        int idx = obj->lm[0] + tile_id_v(obj, pos);                  /* :210-211 */
        v3 n = ld(obj->n);

        double roulette = g->roulette_next;            /* PHILOX: drawn by the previous event */
        rng_event(g, (uint32_t)depth + 1);
        if (pos.z < 0.0005 &&                                        /* :228 mirror, unattenuated */
            (g->mode != ORC_RNG_PHILOX ? rng_u01(g, 0) : roulette) < 0.75) {
            dir = vsub(dir, vmul(n, 2 * vdot(n, dir)));              /* :230 */
            st->mirror_bounces++;
        } else {
            dir = sample_hemisphere(g, n, 0, 0);                     /* :233 */
            if (pos.z < 1E-5f) {                                     /* :236-246 floor tint */
                colour.x *= 1.0f; colour.y *= 0.85f; colour.z *= 0.7f;
            }
            colour = vmul(colour, 0.9f);                             /* :247 */
        }
        if (texels) {                                                /* :251 (lane 3 := 0, add() zero-fills) */
            float *t = texels + 4 * (size_t)idx;
            t[0] = t[0] + colour.x; t[1] = t[1] + colour.y; t[2] = t[2] + colour.z; t[3] = 0;
        }
        if (path) path[depth] = idx;
        st->deposits++;
        pos = vadd(pos, vmul(dir, 1E-5f));                           /* :254 */
    }
}

/* photonmap.c:408-434 */
void orc_bake(const orc_rect *walls, int num_walls, const orc_rect *windows, int num_windows,
              const orc_rect *lights, int num_lights, float *texels, int spa, int max_depth,
              int accel, int rng, uint32_t seed, int shard, int num_shards, orc_stats *stats)
{
    scene_t s = {walls, num_walls, accel, NULL};
    if (accel == ORC_ACCEL_BSP)
        s.root = bsp_build(walls, num_walls);
    orc_stats st;
    memset(&st, 0, sizeof st);
    rng_t g;
    memset(&g, 0, sizeof g);
    g.mode = rng;
    g.seq = 0x1234567ull * (seed + 1);
    if (num_shards < 1) { num_shards = 1; shard = 0; }

    for (int e = 0; e < num_windows + num_lights; e++) {
        int is_window = e < num_windows;
        const orc_rect *src = is_window ? &windows[e] : &lights[e - num_windows];
        uint64_t n = orc_photon_budget(src, spa);
        uint64_t first = 0, last = n;
        if (rng == ORC_RNG_PHILOX) {      /* contiguous photon-index ranges per shard */
            first = n * (uint64_t)shard / (uint64_t)num_shards;
            last = n * (uint64_t)(shard + 1) / (uint64_t)num_shards;
        }
        g.key[0] = seed; g.key[1] = (uint32_t)e;
        for (uint64_t i = first; i < last; i++) {
            g.photon_lo = (uint32_t)i; g.photon_hi = (uint32_t)(i >> 32);
            trace_photon(&s, src, is_window, &g, max_depth, texels, &st, NULL);
        }
    }
    bsp_free(s.root);
    if (stats) *stats = st;
}

void orc_trace_paths(const orc_rect *walls, int num_walls, const orc_rect *emitter, int is_window,
                     int emitter_index, int max_depth, uint32_t seed, uint64_t first, int count,
                     int32_t *texel_out)
{
    scene_t s = {walls, num_walls, ORC_ACCEL_LINEAR, NULL};
    orc_stats st;
    memset(&st, 0, sizeof st);
    rng_t g;
    memset(&g, 0, sizeof g);
    g.mode = ORC_RNG_PHILOX;
    g.key[0] = seed; g.key[1] = (uint32_t)emitter_index;
    for (int i = 0; i < count; i++) {
        uint64_t p = first + (uint64_t)i;
        int32_t *path = texel_out + (size_t)i * max_depth;
        for (int b = 0; b < max_depth; b++) path[b] = -1;
        g.photon_lo = (uint32_t)p; g.photon_hi = (uint32_t)(p >> 32);
        trace_photon(&s, emitter, is_window, &g, max_depth, NULL, &st, path);
    }
}

void orc_closest_hit(const orc_rect *walls, int num_walls, int accel, const float *origins,
                     const float *dirs, int num_rays, int32_t *hit_index, float *hit_dist)
{
    scene_t s = {walls, num_walls, accel, NULL};
    if (accel == ORC_ACCEL_BSP)
        s.root = bsp_build(walls, num_walls);
    uint64_t tests = 0;
    for (int r = 0; r < num_rays; r++) {
        float dist; int hit;
        closest(&s, ld(origins + 3 * r), ld(dirs + 3 * r), &dist, &hit, &tests);
        hit_index[r] = hit;
        hit_dist[r] = dist;
    }
    bsp_free(s.root);
}

/* ------------------------------------------------------------------------------------------
 * main.c:68-79 + rectangle.c:263-336 — normalise, tone-map and pack one wall's base-level tile.
 * Every implicit promotion of the reference is kept: `0.35 * tilesPerSample` is a double product
 * narrowed to float by mul()'s parameter; the luminance sum and exp() run in double; the floor
 * tint multiplies a uint8 by a double and truncates.
 * ---------------------------------------------------------------------------------------- */
static uint8_t clamp_u8(float d)                                     /* rectangle.c:287-292 */
{
    if (d < 0) d = 0;
    if (d > 255) d = 255;
    return d;
}

void orc_tonemap_tiles(const orc_rect *walls, int num_walls, const float *texels, int spa, int tint_extra,
                       uint8_t *rgb_out)
{
    for (int i = 0; i < num_walls; i++) {
        const orc_rect *obj = &walls[i];
        const int base = obj->lm[0], tiles = obj->lm[1] * obj->lm[2];
        float area = vlen(ld(obj->width)) * vlen(ld(obj->height));                 /* rectangle.c:194-197 */
        float tiles_per_sample = tiles / (area * spa);                              /* main.c:73 */
        float scale = 0.35 * tiles_per_sample;                                      /* main.c:77, narrowed by mul() */
        const int is_floor = obj->pos[2] == 0 && obj->width[2] == 0 && obj->height[2] == 0;   /* rectangle.c:317 */
        for (int j = 0; j < tiles; j++) {
            const float *t = texels + 4 * (size_t)(base + j);
            float r = t[0] * scale, g = t[1] * scale, b = t[2] * scale;            /* main.c:77 */
            float luminance = 0.2126 * r + 0.7152 * g + 0.0722 * b;                /* rectangle.c:277 */
            float perceptive = 1 - exp(-2 * luminance);                             /* rectangle.c:269 */
            r *= perceptive / luminance; g *= perceptive / luminance; b *= perceptive / luminance;
            uint8_t *d = rgb_out + 3 * (size_t)j;
            d[0] = clamp_u8(r * 255); d[1] = clamp_u8(g * 255); d[2] = clamp_u8(b * 255);   /* rectangle.c:307-309 */
            if (is_floor) {                                                         /* rectangle.c:317-334 */
                d[1] *= 0.95; d[2] *= 0.9;
                if (tint_extra) { d[0] *= 1.0f; d[1] *= 0.95f; d[2] *= 0.9f; }
            }
        }
        rgb_out += 3 * (size_t)tiles;
    }
}

/* ------------------------------------------------------------------------------------------
 * photonmap.c:436-491 — ambient occlusion; rectangle.c:140-153 (getTileCenter);
 * photonmap.c:31-48 (transformToOrthoNormalBase); vector3_cl.c:152-170 (createBase).
 * ---------------------------------------------------------------------------------------- */
void orc_ambient_occlusion(const orc_rect *walls, int num_walls, float *texels, const float *dirs, int num_dirs,
                           int accel)
{
    scene_t s = {walls, num_walls, accel, NULL};
    if (accel == ORC_ACCEL_BSP)
        s.root = bsp_build(walls, num_walls);
    uint64_t tests = 0;
    for (int i = 0; i < num_walls; i++) {
        const orc_rect *wall = &walls[i];
        v3 n = ld(wall->n);
        /* createBase: c1 starts as z (or y), c2 = normalized(c1 x n), c1 = normalized(c2 x n) */
        v3 b1 = V(0, 0, 1);
        if (fabs(vdot(n, b1)) >= 0.999999f) b1 = V(0, 1, 0);
        v3 b2 = vnormalized(vcross(b1, n));
        b1 = vnormalized(vcross(b2, n));
        const int tw = wall->lm[1], th = wall->lm[2];
        v3 vw = vdiv(ld(wall->width), tw), vh = vdiv(ld(wall->height), th);        /* rectangle.c:144-145 */
        for (int j = 0; j < tw * th; j++) {
            float dist_sum = 0, fac_sum = 0;
            const int tx = j % tw, ty = j / tw;
            for (int k = 0; k < num_dirs; k++) {
                v3 g = ld(dirs + 3 * k);
                float fac = g.z;
                v3 dir = V(g.x * b1.x + g.y * b2.x + g.z * n.x,                     /* photonmap.c:41-45 */
                           g.x * b1.y + g.y * b2.y + g.z * n.y,
                           g.x * b1.z + g.y * b2.z + g.z * n.z);
                v3 pos = vadd3(ld(wall->pos), vmul(vw, tx + 0.5), vmul(vh, ty + 0.5));   /* rectangle.c:150 */
                pos = vadd(pos, vmul(dir, 1E-5));                                   /* photonmap.c:457 */
                float dist; int hit;
                closest(&s, pos, dir, &dist, &hit, &tests);
                if (hit < 0) dist = 10;                                             /* photonmap.c:462-466 */
                dist_sum += dist * fac;
                fac_sum += fac;
            }
            dist_sum /= (fac_sum * 1.5);                                            /* photonmap.c:473 */
            float *t = texels + 4 * (size_t)(wall->lm[0] + j);
            t[0] = dist_sum; t[1] = dist_sum; t[2] = dist_sum; t[3] = 0;
        }
    }
    bsp_free(s.root);
}
