/* TEST INFRASTRUCTURE — CPU restatement of the reference's photon-mapping path.
 *
 * This is the checker for the CUDA path, not part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors for this path
 * (SURVEY.md §4), so the pin is the reference itself run here: with rng = ORC_RNG_LIBC and
 * accel = ORC_ACCEL_BSP this restatement reproduces performPhotonMappingNative
 * (photonmap.c:408) BIT FOR BIT for the same srand() seed (tests/test_oracle_vs_reference.py
 * compares the raw atlases of oracle/_ref/libfmgi_ref.so and of this file), and the committed
 * fixtures under tests/golden/ were produced by the compiled reference
 * (oracle/make_golden.py).
 */
#ifndef FMGI_PHOTON_ORACLE_H
#define FMGI_PHOTON_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to the reference's Rectangle (rectangle.h:19-26): four float4 + one int4,
 * 80 bytes, 16-byte aligned.  lm = {atlas base index, tiles across width, tiles across height, 0}. */
typedef struct __attribute__((aligned(16))) orc_rect {
    float pos[4], width[4], height[4], n[4];
    int32_t lm[4];
} orc_rect;

/* ORC_RNG_SEQ31: a sequential splitmix64 stream drawn in the reference's order and resolution
 * (31-bit value / 2147483647, like rand()/(double)RAND_MAX) - isolates effects of glibc's additive
 * lagged-Fibonacci rand() from effects of the draw order. */
enum { ORC_RNG_LIBC = 0, ORC_RNG_PHILOX = 1, ORC_RNG_SEQ31 = 2 };
enum { ORC_ACCEL_BSP = 0, ORC_ACCEL_LINEAR = 1 };

typedef struct orc_stats {
    uint64_t photons, rays, deposits, mirror_bounces, rect_tests;
} orc_stats;

/* rectangle.c:67-95 */
float orc_intersects(const orc_rect *rect, const float src[3], const float dir[3], float closest);
/* rectangle.c:205-230 */
int orc_tile_id(const orc_rect *rect, const float p[3]);
/* photonmap.c:414-418: (uint64)(int spa * float area) */
uint64_t orc_photon_budget(const orc_rect *emitter, int samples_per_area);

/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 constants). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]);

/* Whole bake, same emitter order and budgets as performPhotonMappingNative (photonmap.c:408-434).
 * texels: numTexels x float4, accumulated in place.
 * rng = ORC_RNG_LIBC   : libc rand() in the reference's draw order; caller seeds with srand().
 * rng = ORC_RNG_PHILOX : the CUDA path's stream (key = {seed, emitter}, counter =
 *                        {photon lo, photon hi, event, 0}; word layout in photon_oracle.c).
 * shard/num_shards select the contiguous sub-range [N*shard/num_shards, N*(shard+1)/num_shards)
 * of every emitter's N photons in PHILOX mode, mirroring the multi-GPU sharding; ignored in
 * LIBC mode (one sequential rand() stream cannot be split). */
void orc_bake(const orc_rect *walls, int num_walls,
              const orc_rect *windows, int num_windows,
              const orc_rect *lights, int num_lights,
              float *texels, int samples_per_area, int max_depth,
              int accel, int rng, uint32_t seed,
              int shard, int num_shards, orc_stats *stats);

/* Per-photon path record in PHILOX mode: for photons [first, first+count) of emitter
 * `emitter_index`, texel_out[(i*max_depth)+b] = atlas index deposited at bounce b, or -1. */
void orc_trace_paths(const orc_rect *walls, int num_walls,
                     const orc_rect *emitter, int is_window, int emitter_index,
                     int max_depth, uint32_t seed, uint64_t first, int count, int32_t *texel_out);

/* Batch closest hit (linear scan or BSP): hit index into walls (-1 = miss) and distance. */
void orc_closest_hit(const orc_rect *walls, int num_walls, int accel,
                     const float *origins, const float *dirs, int num_rays,
                     int32_t *hit_index, float *hit_dist);

/* performAmbientOcclusionNative (photonmap.c:436-491): every base-level texel of every wall is
 * OVERWRITTEN with (d, d, d, 0), d = sum_k dist_k * fac_k / (1.5 * sum_k fac_k) over the direction
 * set dirs (xyz triples in the texel's local frame, fac = z; the reference uses geoSphere4), where
 * dist_k is the closest-hit distance from the texel centre or 10 for a miss. */
void orc_ambient_occlusion(const orc_rect *walls, int num_walls, float *texels, const float *dirs, int num_dirs,
                           int accel);

/* The caller-side post-processing the lightmap goes through before it becomes tiles/tile_N.png:
 * main.c:68-79 (texel *= 0.35 * tiles / (area * samplesPerArea), base level only) followed by
 * saveAs_core (rectangle.c:293-336: tone-map 1 - exp(-2 L) at constant chroma, x255, clamp, floor
 * tint).  texels: RAW sums (numTexels x float4, not modified).  rgb_out: for every wall in order,
 * tilesW * tilesH RGB bytes - the buffer saveAs hands to write_png_file. */
void orc_tonemap_tiles(const orc_rect *walls, int num_walls, const float *texels, int samples_per_area,
                       int tint_extra, uint8_t *rgb_out);

#ifdef __cplusplus
}
#endif
#endif
