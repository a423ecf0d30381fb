/* TEST INFRASTRUCTURE — thin driver around the UNMODIFIED reference sources.
 *
 * oracle/Makefile compiles this file together with the reference's own
 * photonmap.c, rectangle.c, vector3_cl.c, geometry.c, parseLayout.c, image.c,
 * helpers.c and geoSphere.c (taken where they lie under /root/reference, never
 * copied) into oracle/_ref/libfmgi_ref.so.  Nothing in the product path links or
 * loads that library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * What this file adds on top of the reference:
 *   - the two libpng entry points image.c needs at link time (png_helper.h:13-15);
 *     png.h is not installed here, so layouts enter as raw 0xAABBGGRR pixels
 *     (the word format image.c:189-199 produces) through fmgi_ref_parse_rgba();
 *   - a seeded, timed wrapper around performPhotonMappingNative (photonmap.c:408);
 *   - batch wrappers around intersects (rectangle.c:67), getTileIdAt (rectangle.c:205),
 *     findClosestIntersection (photonmap.c:54) and the two hemisphere samplers
 *     (vector3_cl.c:102,129) so that parity tests can probe them by pointer
 *     instead of passing 16-byte-aligned unions by value through ctypes.
 *
 * When built with -DFMGI_REF_RUNTIME_DEPTH the translation unit of photonmap.c is
 * fed through sed (see Makefile) so that its function-local `MAX_DEPTH = 8`
 * (photonmap.c:173) reads the global below instead; BASELINE.json's 3- and
 * 4-bounce configurations need that and the reference has no knob for it.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "geometry.h"
#include "global_illumination_native.h"
#include "image.h"
#include "parseLayout.h"
#include "png_helper.h"
#include "rectangle.h"
#include "vector3_cl.h"

#ifdef FMGI_REF_RUNTIME_DEPTH
int fmgi_ref_max_depth = 8;
#endif

/* ---- link-time holes left by png_helper.c (needs <png.h>, absent) ------------------ */

void read_png_file(const char *file_name, int *width, int *height, int *color_type,
                   uint8_t **pixel_buffer)
{
    (void)width; (void)height; (void)color_type; (void)pixel_buffer;
    fprintf(stderr, "[fmgi_ref] read_png_file(%s): libpng is not available in the oracle build; "
                    "use fmgi_ref_parse_rgba()\n", file_name);
    exit(2);
}

/* saveAs (rectangle.c:338-346) hands the tone-mapped tile to write_png_file: capture it instead of
 * encoding it, so that tests can read the reference's own pixel bytes. */
static uint8_t *g_capture = NULL;
static size_t g_capture_size = 0;

void write_png_file(const char *file_name, int width, int height, int color_type,
                    uint8_t *pixel_buffer)
{
    /* parseLayout.c:314 also dumps ./filled.png (RGBA) as a debugging side effect; never kept. */
    (void)file_name;
    if (!g_capture || color_type != PNG_COLOR_TYPE_RGB) return;
    size_t n = (size_t)width * height * 3;
    if (n > g_capture_size) n = g_capture_size;
    memcpy(g_capture, pixel_buffer, n);
}

/* The reference's saveAs on one wall of an already normalised atlas; out receives tilesW*tilesH*3 bytes. */
void fmgi_ref_save_tile(const Rectangle *rect, const Vector3 *texels, int tint_extra, uint8_t *out, size_t out_size)
{
    g_capture = out; g_capture_size = out_size;
    saveAs(rect, "capture", texels, tint_extra);
    g_capture = NULL; g_capture_size = 0;
}

/* ---- layout -> Geometry through the reference's own parser ------------------------- */

int fmgi_ref_sizeof_rectangle(void) { return (int)sizeof(Rectangle); }
int fmgi_ref_sizeof_geometry(void)  { return (int)sizeof(Geometry); }
int fmgi_ref_sizeof_vector3(void)   { return (int)sizeof(Vector3); }
int fmgi_ref_runtime_depth(void)
{
#ifdef FMGI_REF_RUNTIME_DEPTH
    return 1;
#else
    return 0;
#endif
}

/* pixels: width*height words 0xAABBGGRR (image.c:189-199); scale in pixels per metre
 * as on the command line (main.c:32,45). */
Geometry *fmgi_ref_parse_rgba(const uint32_t *pixels, int width, int height,
                              float pixels_per_metre, float tile_size)
{
    Image img;
    img.width = width;
    img.height = height;
    img.data = (uint32_t *)pixels;              /* parseLayout clones before mutating */
    return parseLayout(&img, 1 / pixels_per_metre, tile_size);
}

char *fmgi_ref_collision_map_json(const uint32_t *pixels, int width, int height)
{
    Image img;
    img.width = width;
    img.height = height;
    img.data = (uint32_t *)pixels;
    return buildCollisionMap(&img);
}

char *fmgi_ref_geometry_json(Geometry *geo) { return getJsonString(geo); }
void fmgi_ref_free_string(char *s) { free(s); }
void fmgi_ref_free_geometry(Geometry *geo) { freeGeometry(geo); }

/* ---- the oracle call itself -------------------------------------------------------- */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* srand(seed) then performPhotonMappingNative (photonmap.c:408); returns wall seconds.
 * depth is honoured only by the FMGI_REF_RUNTIME_DEPTH build (else it must be 8). */
double fmgi_ref_photonmap_native(Geometry *geo, int samples_per_area, unsigned seed, int depth)
{
#ifdef FMGI_REF_RUNTIME_DEPTH
    fmgi_ref_max_depth = depth;
#else
    if (depth != 8) {
        fprintf(stderr, "[fmgi_ref] this build has the reference's fixed MAX_DEPTH=8\n");
        exit(2);
    }
#endif
    srand(seed);
    double t0 = now_s();
    performPhotonMappingNative(geo, samples_per_area);
    return now_s() - t0;
}

/* ---- probes for unit-level parity -------------------------------------------------- */

/* Linear closest-hit scan with the reference's intersects(); strict '<' keeps the
 * lowest index on ties, as the OpenCL kernel's loop does (photonmap.cl:194-206). */
void fmgi_ref_closest_hit_linear(const Rectangle *rects, int num_rects,
                                 const float *origins, const float *dirs, int num_rays,
                                 int *hit_index, float *hit_dist)
{
    for (int r = 0; r < num_rays; r++) {
        Vector3 o = vec3(origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]);
        Vector3 d = vec3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        float best = INFINITY;
        int best_i = -1;
        for (int i = 0; i < num_rects; i++) {
            float t = intersects(&rects[i], o, d, best);
            if (t < 0)
                continue;
            if (t < best) {
                best = t;
                best_i = i;
            }
        }
        hit_index[r] = best_i;
        hit_dist[r] = best;
    }
}

/* findClosestIntersection is not declared in any reference header (photonmap.c:54). */
int findClosestIntersection(Vector3 ray_pos, Vector3 ray_dir, const struct BspTreeNode *node,
                            float *dist, float distShift, Rectangle **targetOut, int depth);

/* Closest hit through the reference's BSP.  hit_index is recovered from the atlas base
 * offset (lightmapSetup[0]) of the copied Rectangle the tree hands back. */
void fmgi_ref_closest_hit_bsp(Rectangle *rects, int num_rects,
                              const float *origins, const float *dirs, int num_rays,
                              int *hit_base, float *hit_dist)
{
    struct BspTreeNode *root = buildBspTree(rects, num_rects);
    for (int r = 0; r < num_rays; r++) {
        Vector3 o = vec3(origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]);
        Vector3 d = vec3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        float best = INFINITY;
        Rectangle *target = NULL;
        findClosestIntersection(o, d, root, &best, 0, &target, 0);
        hit_base[r] = target ? target->lightmapSetup.s[0] : -1;
        hit_dist[r] = best;
    }
    freeBspTree(root);
}

void fmgi_ref_tile_ids(const Rectangle *rect, const float *points, int num_points, int *tile_ids)
{
    for (int i = 0; i < num_points; i++)
        tile_ids[i] = getTileIdAt(rect, vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
}

/* n draws of the reference samplers for one normal; libc rand() state is the caller's. */
void fmgi_ref_sample_dirs(const float *normal, int sky, unsigned seed, int n, float *out)
{
    Vector3 nd = vec3(normal[0], normal[1], normal[2]);
    srand(seed);
    for (int i = 0; i < n; i++) {
        Vector3 d = sky ? getDiffuseSkyRandomRay(nd) : getCosineDistributedRandomRay(nd);
        out[3 * i] = d.s[0];
        out[3 * i + 1] = d.s[1];
        out[3 * i + 2] = d.s[2];
    }
}
