/* fmgi.h — C ABI of the B200 photon-mapping lightmap baker (libfmgi_cuda.so).
 *
 * Drop-in boundary: the library exports the reference's own entry point
 *
 *     void performGlobalIlluminationCl(Geometry *geo, int numSamplesPerArea);
 *
 * declared by the reference in global_illumination_cl.h:10 and called once from main.c:63.
 * It replaces global_illumination_cl.c (OpenCL context/program/buffer/enqueue code, :148-321)
 * and photonmap.cl (the kernel, :269).  The reference's main.c, parseLayout.c, geometry.c and
 * the tiles/, geometry.json and collisionMap.json writers are compiled untouched and link
 * against this library instead of global_illumination_cl.o and -lOpenCL.
 *
 * Everything else in this header is our extension for harnesses (depth, seed, sharding,
 * device-resident atlases, counters) and for unit-level parity probes.  Only plain C types,
 * pointers and sizes cross the boundary.
 *
 * Error convention: the reference's entry point returns void and exit()s on failure
 * (global_illumination_cl.c:208-209,229-230,241,254,263); performGlobalIlluminationCl keeps
 * that (prints "[Err] ..." and exit(1)).  The fmgi_* functions return 0 on success and a
 * negative fmgi_status otherwise; fmgi_last_error() gives the message.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with FMGI_ERR_CUDA.
 */
#ifndef FMGI_H
#define FMGI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- data layout shared with the reference ------------------------------------------------- */

/* Layout-identical to the reference's Rectangle (rectangle.h:19-26): pos, width, height, n as
 * cl_float4 (lane 3 unused, zero) and lightmapSetup as cl_int4 =
 * {atlas base index, tiles along width, tiles along height, 0}.  80 bytes, 16-byte aligned. */
typedef struct __attribute__((aligned(16))) fmgi_rect {
    float   pos[4];
    float   width[4];
    float   height[4];
    float   n[4];
    int32_t lightmap[4];
} fmgi_rect;

/* Layout-identical to the reference's Geometry (geometry.h:7-15); 80 bytes on x86-64.
 * texels is numTexels x float4 (Vector3 = cl_float4, vector3_cl.h:14), 16-byte stride. */
typedef struct fmgi_geometry {
    fmgi_rect *windows, *lights, *walls, *boxWalls;
    int32_t    numWindows, numLights, numWalls, numBoxWalls;
    int32_t    width, height;
    float      startingPositionX, startingPositionY;
    int32_t    numTexels;
    float     *texels;
} fmgi_geometry;

struct Geometry;   /* the reference's own type; same bytes as fmgi_geometry */

/* ---- the reference boundary (global_illumination_cl.h:10) ---------------------------------- */

/* geo->texels += raw photon deposits of all windows then all lights (photonmap.c:408-431 budget:
 * N = (uint64)(numSamplesPerArea * area) photons per emitter), path length <= 8
 * (photonmap.c:173).  Lane 3 of every texel and the mip-chain slots are left untouched;
 * normalisation stays with the caller (main.c:68-79).  Environment overrides for harnesses:
 * FMGI_MAX_DEPTH, FMGI_SEED, FMGI_GPUS, FMGI_STATS=1 (print counters). */
void performGlobalIlluminationCl(struct Geometry *geo, int numSamplesPerArea);

/* ---- extension: options, counters ---------------------------------------------------------- */

typedef enum fmgi_status {
    FMGI_OK = 0,
    FMGI_ERR_CUDA = -1,       /* no device / CUDA runtime error */
    FMGI_ERR_ARG = -2,        /* bad argument */
    FMGI_ERR_UNSUPPORTED = -3 /* scene exceeds what the selected tier supports */
} fmgi_status;

enum { FMGI_DEPOSIT_VEC4 = 0, FMGI_DEPOSIT_SCALAR = 1, FMGI_DEPOSIT_WARP_AGG = 2 };
enum { FMGI_TIER_AUTO = 0, FMGI_TIER_SOUP = 1, FMGI_TIER_GRID = 2, FMGI_TIER_ROOMS = 4 };

typedef struct fmgi_options {
    uint32_t struct_size;     /* = sizeof(fmgi_options); lets the struct grow */
    int32_t  max_depth;       /* bounces per photon, 1..15; reference = 8 (photonmap.c:173) */
    uint32_t seed;            /* Philox key */
    int32_t  num_gpus;        /* fmgi_bake only: GPUs to shard photons over (0 = 1) */
    int32_t  shard;           /* this caller's shard of every emitter's photon range ... */
    int32_t  num_shards;      /* ... out of num_shards (0/1 = everything) */
    int32_t  tier;            /* FMGI_TIER_* */
    int32_t  deposit;         /* FMGI_DEPOSIT_* */
    int32_t  device;          /* CUDA device ordinal for the fmgi_scene_* calls */
    int32_t  count_tests;     /* != 0: also count the rectangle tests of the grid lookups (fmgi_stats.rect_tests);
                                 costs two instructions per walk step, so off by default */
    int32_t  reserved[6];
} fmgi_options;

typedef struct fmgi_stats {
    uint64_t photons;         /* photons emitted */
    uint64_t rays;            /* closest-hit queries */
    uint64_t deposits;        /* texel deposits = photon-bounces (the BASELINE metric's unit) */
    uint64_t mirror_bounces;
    uint64_t rect_tests;      /* rectangle tests executed (all lanes); grid lookups only with count_tests; room tier
                                 (count_tests): boxes crossed + face grids looked up */
    uint64_t kernel_launches; /* our kernels launched */
    double   trace_ms;        /* device time of the trace kernels (CUDA events), max over GPUs */
    double   h2d_ms, d2h_ms;  /* atlas upload (runs under the trace) / read-back, CUDA events, max over GPUs */
    double   reduce_ms;       /* atlas fold: reduce-scatter over peer memory + the caller's values (CUDA events) */
    double   total_ms;        /* whole call, host clock */
    int32_t  num_gpus;
    int32_t  tier;
    int32_t  num_sms;
    int32_t  sm_clock_khz;    /* cudaDevAttrClockRate */
    /* host-buffer entry points only (fmgi_bake, fmgi_bake_tiles), host clock, max over GPUs */
    double   init_ms;         /* cudaSetDevice + stream/event creation: the CUDA context on the first call of a process */
    double   prepare_ms;      /* rectangle tables (scene_prep.cpp) */
    double   grid_build_ms;   /* floor-plan grid: host classification + per-cell assembly (host, or device for >= 2048 colliders);
                                 room tier: the box decomposition (rooms_build.cpp, host thread pool) */
    double   upload_ms;       /* table upload + kernel attribute queries per GPU */
    int32_t  pool_rays;       /* 0: k_trace ran; K > 0: the pooled kernel (trace_pool.cuh) with K rays per lane */
    int32_t  bounds_violations; /* -1: regular build; >= 0: lib/libfmgi_cuda_checked.so (-DFMGI_CHECKED) - data-dependent
                                 indices (grid records, shading records, texels) found outside their tables */
} fmgi_stats;

void        fmgi_default_options(fmgi_options *opt);
const char *fmgi_last_error(void);
const char *fmgi_version(void);
/* sha256 prefix of the sources the library was built from (Makefile); profiles/ncu_facts.json carries the hash of the
 * build its ncu counters were captured on, and bench.py marks them stale when the two differ. */
const char *fmgi_source_hash(void);
int         fmgi_device_count(void);
/* The library keeps freed device / pinned blocks in a process-wide cache so that repeated bakes do
 * not pay cudaMalloc/cudaFree again: at most 256 MB after performGlobalIlluminationCl (atlas-sized blocks are
 * released before it returns), at most FMGI_CACHE_MB (default 8192) after fmgi_bake / fmgi_bake_tiles.
 * fmgi_release_cache() returns the cached blocks to the driver; fmgi_cached_bytes() reports them. */
void        fmgi_release_cache(void);
uint64_t    fmgi_cached_bytes(void);

/* Host-buffer bake with options: what performGlobalIlluminationCl wraps. */
int fmgi_bake(struct Geometry *geo, int numSamplesPerArea, const fmgi_options *opt, fmgi_stats *stats);

/* ---- extension: device-resident scenes (inputs already in HBM) ----------------------------- */

typedef struct fmgi_scene fmgi_scene;

/* Uploads the collider table (geo->walls only collide, photonmap.c:410) and the emitter table
 * (windows then lights) to opt->device and precomputes the traversal tables. */
int  fmgi_scene_create(fmgi_scene **out, const fmgi_rect *walls, int num_walls,
                       const fmgi_rect *windows, int num_windows,
                       const fmgi_rect *lights, int num_lights,
                       int num_texels, const fmgi_options *opt);
void fmgi_scene_destroy(fmgi_scene *scene);

/* Enqueues the whole bake (all emitters, this shard) on `cuda_stream` (a cudaStream_t, or NULL
 * for the default stream), accumulating into the device atlas `atlas_dev`
 * (numTexels x float4).  Asynchronous; counters become valid after fmgi_scene_sync(). */
int fmgi_scene_trace(fmgi_scene *scene, void *atlas_dev, int numSamplesPerArea,
                     const fmgi_options *opt, void *cuda_stream);
/* Waits for the stream used by the last fmgi_scene_trace and returns its counters;
 * trace_ms is the CUDA-event time of the kernels enqueued by that call. */
int fmgi_scene_sync(fmgi_scene *scene, fmgi_stats *stats);
/* Photons this shard emits for the given density (sum over emitters). */
uint64_t fmgi_scene_photon_count(const fmgi_scene *scene, int numSamplesPerArea, const fmgi_options *opt);

/* ---- extension: tile post-processing on the device ------------------------- */

/* Bytes of the packed tile buffer: sum over walls of tilesW * tilesH * 3. */
uint64_t fmgi_tile_bytes(const fmgi_rect *walls, int num_walls);
/* Device version of main.c:68-79 (normalisation) + saveAs_core (rectangle.c:293-336): from the RAW
 * device atlas to, for every wall in order, tilesW * tilesH RGB bytes - the pixel buffer saveAs
 * hands to write_png_file.  rgb_dev: device buffer of fmgi_tile_bytes() bytes.  Asynchronous on
 * `cuda_stream`. */
int fmgi_scene_tonemap(fmgi_scene *scene, const void *atlas_dev, int numSamplesPerArea, int tintExtra,
                       void *rgb_dev, void *cuda_stream);
/* fmgi_bake followed by the tone-map on the device: reads back 3 bytes per texel instead of 16.
 * geo->texels is the initial atlas and is NOT written back; rgb_out receives fmgi_tile_bytes() bytes. */
int fmgi_bake_tiles(struct Geometry *geo, int numSamplesPerArea, const fmgi_options *opt, int tintExtra,
                    uint8_t *rgb_out, fmgi_stats *stats);

/* The same tiles as complete PNG FILES assembled on the device (rectangle.c:338-346 saveAs + png_helper.c:255
 * write_png_file): 8-bit RGB, filter 0, zlib stream of stored blocks (no compression, so the layout is known in
 * advance), Adler-32 and CRC-32 computed on the GPU.  Wall i's file is bytes [offsets[i], offsets[i + 1]) of the
 * output; fmgi_tile_png_bytes returns the total and fills offsets_out (num_walls + 1 entries, may be NULL).  A decoder
 * reads the pixels write_png_file would have written. */
uint64_t fmgi_tile_png_bytes(const fmgi_rect *walls, int num_walls, uint64_t *offsets_out);
int fmgi_scene_tiles_png(fmgi_scene *scene, const void *atlas_dev, int numSamplesPerArea, int tintExtra,
                         void *png_dev, void *cuda_stream);
int fmgi_bake_tiles_png(struct Geometry *geo, int numSamplesPerArea, const fmgi_options *opt, int tintExtra,
                        uint8_t *png_out, fmgi_stats *stats);

/* ---- extension: ambient occlusion on the device ---------------------------- */

/* performAmbientOcclusionNative (global_illumination_native.h:16, photonmap.c:436-491) on the closest-hit
 * code of the photon tracer: every base-level texel of every wall is OVERWRITTEN with (d, d, d, 0),
 * d = sum(dist * fac) / (1.5 * sum(fac)) over the reference's 481 geoSphere4 directions (dist = 10 for
 * a miss).  Other texels keep their contents. */
int fmgi_ambient_occlusion(struct Geometry *geo, const fmgi_options *opt);
/* The direction set itself (host only, no GPU needed): geodesic half-sphere with `iterations`
 * subdivisions; 4 gives the reference's geoSphere4 (geoSphere.c:148) as a set.  Returns the count. */
int fmgi_geosphere(int iterations, float *xyz_out, int max_directions);
/* The same on a device-resident atlas; asynchronous on `cuda_stream` after its table upload. */
int fmgi_scene_ambient_occlusion(fmgi_scene *scene, void *atlas_dev, void *cuda_stream);

/* ---- extension: parity probes (each runs the same device functions the trace kernel uses) -- */

/* Closest front-facing hit for num_rays host rays (xyz triples): wall index or -1, distance. */
int fmgi_probe_closest_hit(fmgi_scene *scene, const float *origins, const float *dirs, int num_rays,
                           int32_t *hit_index, float *hit_dist);
/* getTileIdAt (rectangle.c:205) for points on walls[rect_index[i]]. */
int fmgi_probe_tile_ids(fmgi_scene *scene, const int32_t *rect_index, const float *points,
                        int num_points, int32_t *tile_ids);
/* The floor-plan grid table T the trace kernels walk (32-byte records, csrc/scene_tables.h), as it sits on the device:
 * copies up to max_records records to records_out (may be NULL) and returns the table's record count, -1 on error.
 * Scenes of 2048 colliders and more assemble T on the GPU (csrc/grid_build.cuh; FMGI_GRID_BUILD=host|device
 * overrides); the probe lets a test compare the two builders bit for bit. */
int64_t fmgi_probe_grid_table(fmgi_scene *scene, void *records_out, uint64_t max_records);
/* Philox4x32-10 block on the device (known-answer probe; the tracer itself draws Philox2x32-10 blocks). */
int fmgi_probe_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Philox2x32-10 block on the device: the tracer's generator (csrc/philox.cuh: key = seed, counter =
 * {photon lo, photon hi | emitter << 8 | event << 28}). */
int fmgi_probe_philox2x32(const uint32_t ctr[2], uint32_t key, uint32_t out[2]);
/* Deposit roofline: rate of the trace kernel's deposit instruction (one 16-byte vector
 * reduction, RED.E.ADD.F32x4) at uniform-random texels of a scratch atlas of num_texels float4, nothing else
 * in the loop.  Deposits per second, device-timed. */
int fmgi_probe_deposit_peak(uint64_t num_texels, uint64_t num_deposits, int device, double *deposits_per_s);
/* n directions around `normal` from the device sampler (sky != 0: window fold). */
int fmgi_probe_sample_dirs(const float normal[3], int sky, uint32_t seed, int n, float *dirs_out);
/* Per-photon paths: atlas index deposited at bounce b of photons [first, first+count) of
 * emitter `emitter_index` (windows first), -1 where the photon is dead. */
int fmgi_probe_paths(fmgi_scene *scene, int emitter_index, int max_depth, uint32_t seed,
                     uint64_t first, int count, int32_t *texel_out);

#ifdef __cplusplus
}
#endif
#endif /* FMGI_H */
