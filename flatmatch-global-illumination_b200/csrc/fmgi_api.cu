// C ABI of libfmgi_cuda.so (include/fmgi.h): the reference's performGlobalIlluminationCl entry
// point (global_illumination_cl.h:10), the options/counters extension and the parity probes.
// Replaces the OpenCL host code of global_illumination_cl.c:148-321.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mem_pool.h"
#include "grid_build.cuh"
#include "room_tables.h"
#include "trace_core.cuh"
#include "trace_pool.cuh"
#include "tile_kernels.cuh"

using namespace fmgi;

namespace {

// photons per fp32 accumulation pass (see fmgi_scene_trace)
constexpr unsigned long long kAccumPhotons = 1ull << 28;

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define FMGI_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t err__ = (expr);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            return fail(FMGI_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));      \
    } while (0)

// Wall time since the process was started, from /proc/self/stat (field 22, clock ticks since boot) and
// /proc/uptime; 10 ms resolution.  Only for the FMGI_STATS=2 breakdown line.
double ms_since_process_start()
{
    FILE *f = fopen("/proc/self/stat", "r");
    if (!f) return 0.0;
    char buf[2048];
    const size_t n = fread(buf, 1, sizeof buf - 1, f);
    fclose(f);
    buf[n] = 0;
    const char *p = strrchr(buf, ')');            // the command name may contain spaces
    if (!p) return 0.0;
    unsigned long long start = 0;
    int field = 2;
    for (p++; *p && field < 22; p++)
        if (*p == ' ' && ++field == 22) sscanf(p + 1, "%llu", &start);
    double up = 0.0;
    f = fopen("/proc/uptime", "r");
    if (!f) return 0.0;
    if (fscanf(f, "%lf", &up) != 1) up = 0.0;
    fclose(f);
    const double hz = (double)sysconf(_SC_CLK_TCK);
    return up > 0.0 && hz > 0.0 ? (up - (double)start / hz) * 1e3 : 0.0;
}

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

fmgi_options resolve(const fmgi_options *in)
{
    fmgi_options o;
    fmgi_default_options(&o);
    if (in) {
        size_t n = in->struct_size ? in->struct_size : sizeof(fmgi_options);
        if (n > sizeof(fmgi_options)) n = sizeof(fmgi_options);
        memcpy(&o, in, n);
        o.struct_size = sizeof(fmgi_options);
    }
    if (o.max_depth <= 0) o.max_depth = 8;
    if (o.num_shards <= 0) { o.num_shards = 1; o.shard = 0; }
    if (o.num_gpus <= 0) o.num_gpus = 1;
    return o;
}

// Pool-backed device buffer that goes back to the pool on every exit path.
template <typename T>
struct DevBuf {
    T *ptr = nullptr;
    cudaError_t alloc(size_t count) { return MemPool::get().alloc((void **)&ptr, (count ? count : 1) * sizeof(T), false); }
    ~DevBuf() { MemPool::get().free(ptr); }
    operator T *() const { return ptr; }
};

template <typename T>
cudaError_t upload(T **dst, const std::vector<T> &src)
{
    *dst = nullptr;
    size_t bytes = src.size() * sizeof(T);
    cudaError_t e = MemPool::get().alloc((void **)dst, bytes ? bytes : sizeof(T), false);
    if (e != cudaSuccess) return e;
    if (bytes) e = cudaMemcpy(*dst, src.data(), bytes, cudaMemcpyHostToDevice);
    return e;
}

}  // namespace

// Host-side precompute of one scene (scene_prep.cpp): built once per bake, uploaded to every GPU that takes part.
struct HostBuild {
    HostScene scene;
    std::vector<float> wall_wh;                 // the walls' width and height vectors (6 floats per wall)
    std::vector<float> wall_area;               // |width| * |height| per wall, float (rectangle.c:194-197)
    std::vector<int> wall_floor;                // rectangle.c:317
    int tier = FMGI_TIER_SOUP;
    int kernel_tier = FMGI_TIER_SOUP;           // tier, or kTierSoupPlanes (soup + the grid's plane tables)
    size_t smem_bytes = 0;
    uint64_t tests_per_ray = 0;
    double prepare_ms = 0, grid_ms = 0;         // host clock: rectangle tables, floor-plan grid
    RoomScene rooms;                            // room tier: box decomposition (rooms_build.cpp)
    double rooms_ms = 0;
    std::vector<GridItem> grid_items;           // classified colliders, kept when T is assembled on the device
    bool device_grid = false;                   // T is assembled per GPU by grid_build.cuh instead of on the host
};

struct fmgi_scene {
    int device = 0;
    std::shared_ptr<HostBuild> build;           // host-side tables, shared by the per-GPU scenes of one bake
    HostScene &host;
    explicit fmgi_scene(std::shared_ptr<HostBuild> b) : build(std::move(b)), host(build->scene) {}
    // device tables
    AxisPairBlock *d_axis = nullptr;
    GeneralRect *d_general = nullptr;
    ShadeRect *d_shade = nullptr;
    EmitterRec *d_emitters = nullptr;
    GridRec *d_grid_table = nullptr;            // grid tier / plane tables
    RoomBox *d_room_boxes = nullptr;            // room tier
    RoomFaceGrid *d_room_face_grids = nullptr;
    uint32_t *d_room_face_cells = nullptr;
    RoomBounds *d_room_bounds = nullptr;
    RoomNode *d_room_nodes = nullptr;
    RoomStart *d_room_starts = nullptr;
    unsigned long long *d_jobs = nullptr;       // per accumulation pass: chunk_begin[E+1], photon_first[E], photon_count[E]
    size_t job_tables = 0;                      // passes the job-table buffers have room for
    float4 *d_scratch = nullptr;                // per-pass fp32 atlas when a bake needs several passes
    TileWall *d_tile_walls = nullptr;           // tone-map wall table (fmgi_scene_tonemap)
    TileWall *h_tile_walls = nullptr;           // pinned staging for it
    PngWall *d_png_walls = nullptr, *h_png_walls = nullptr;   // PNG layout of the tiles (fmgi_scene_tiles_png)
    float *d_ao = nullptr;                      // ambient occlusion: widths, heights (3 floats per wall), then float4 directions
    AoWall *d_ao_walls = nullptr;
    unsigned long long *d_counters = nullptr;   // 4 counters + work counter
    unsigned long long *h_jobs = nullptr;       // pinned staging
    unsigned long long *h_counters = nullptr;   // pinned
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaStream_t last_stream = nullptr;
    std::vector<cudaStream_t> streams;          // every stream work of this scene was enqueued on
    void note_stream(cudaStream_t st)
    {
        last_stream = st;
        for (cudaStream_t t : streams) if (t == st) return;
        streams.push_back(st);
    }
    bool traced = false;
    int num_sms = 0, clock_khz = 0, blocks_per_sm = 0;
    size_t smem_bytes = 0;
    int tier = FMGI_TIER_SOUP;
    int kernel_tier = FMGI_TIER_SOUP;           // tier, or kTierSoupPlanes (soup + the grid's plane tables)
    uint64_t launches = 0;
    uint64_t tests_per_ray = 0;
    int min_blocks = 4;                         // resident CTAs per SM the trace kernel is compiled for
    int pool_k = 0;                             // pooled kernel (trace_pool.cuh): rays per lane, 0 = not usable for this scene
    int pool_blocks_per_sm = 0;
    bool pooled_last = false;                   // the last trace ran the pooled kernel
    size_t grid_records = 0;                    // records of T on the device
    double grid_device_ms = 0;                  // device assembly of T (grid_build.cuh), host clock incl. its sync
};

namespace {

TraceParams base_params(const fmgi_scene *s)
{
    TraceParams p;
    memset(&p, 0, sizeof p);
    p.axis = reinterpret_cast<const float4 *>(s->d_axis);
    p.general = reinterpret_cast<const float4 *>(s->d_general);
    for (int g = 0; g < 4; g++) p.pair_begin[g] = s->host.pair_begin[g];
    p.num_general = (int)s->host.general.size();
    p.grid_table = reinterpret_cast<const float4 *>(s->d_grid_table);
    p.grid = s->host.grid;
    p.grid_has_misc = s->host.grid_misc != 0;
    p.one = 1;
    p.shade = reinterpret_cast<const float4 *>(s->d_shade);
    p.emitters = reinterpret_cast<const float4 *>(s->d_emitters);
    p.num_emitters = (int)s->host.emitters.size();
    p.job_begin = s->d_jobs;
    p.photon_first = s->d_jobs + (p.num_emitters + 1);
    p.photon_count = s->d_jobs + (2 * p.num_emitters + 1);
    p.work_counter = s->d_counters + 4;
    p.counters = s->d_counters;
    p.room_boxes = reinterpret_cast<const float4 *>(s->d_room_boxes);
    p.room_face_grids = reinterpret_cast<const float4 *>(s->d_room_face_grids);
    p.room_face_cells = s->d_room_face_cells;
    p.room_bounds = reinterpret_cast<const float4 *>(s->d_room_bounds);
    p.room_nodes = reinterpret_cast<const float4 *>(s->d_room_nodes);
    p.room_starts = reinterpret_cast<const int2 *>(s->d_room_starts);
    for (int k = 0; k < 3; k++) { p.room_lo[k] = s->build->rooms.root_lo[k]; p.room_hi[k] = s->build->rooms.root_hi[k]; }
    p.room_num_boxes = (unsigned)s->build->rooms.boxes.size();
    p.room_num_face_grids = (unsigned)s->build->rooms.face_grids.size();
    p.room_num_face_cells = (unsigned)s->build->rooms.face_cells.size();
    p.room_num_nodes = (unsigned)s->build->rooms.nodes.size();
    p.grid_records = (unsigned)s->grid_records;
    p.num_walls = (unsigned)s->host.num_walls;
    p.num_texels = (unsigned)s->host.num_texels;
#ifdef FMGI_CHECKED
    if (getenv("FMGI_CHECK_SELFTEST")) p.num_texels /= 2;      // makes the checks fire: half of the deposits are "outside"
#endif
    return p;
}

void set_seed(TraceParams &p, uint32_t seed)
{
    p.seed = seed;
    for (int r = 0; r < 10; r++) p.philox_keys[r] = seed + (uint32_t)r * kPhiloxW;     // Philox2x32 round keys
}

// Job tables, per accumulation pass: chunk_begin[E + 1] (prefix of the emitters' chunk counts; a chunk is
// `chunk` consecutive photon indices of ONE emitter, the unit a warp claims), photon_first[E],
// photon_count[E].
inline size_t job_table_words(size_t E) { return 3 * E + 2; }

// Fills the pinned job tables for (spa, shard): emitter e's N photons (photonmap.c:414-418) are split
// into num_shards contiguous index ranges.  Returns the shard's photon total; *chunks its chunk total.
unsigned long long fill_jobs(const fmgi_scene *s, int spa, const fmgi_options &o, unsigned long long *jobs,
                             unsigned long long *chunks = nullptr, int chunk = kChunkPhotons)
{
    const int E = (int)s->host.emitters.size();
    unsigned long long total = 0, total_chunks = 0;
    for (int e = 0; e < E; e++) {
        const unsigned long long n = photon_budget(s->host.emitter_area[e], spa);
        const unsigned long long first = (unsigned long long)((unsigned __int128)n * o.shard / o.num_shards);
        const unsigned long long last = (unsigned long long)((unsigned __int128)n * (o.shard + 1) / o.num_shards);
        jobs[e] = total_chunks;
        jobs[E + 1 + e] = first;
        jobs[2 * E + 1 + e] = last - first;
        total += last - first;
        total_chunks += (last - first + chunk - 1) / chunk;
    }
    jobs[E] = total_chunks;
    if (chunks) *chunks = total_chunks;
    return total;
}

// Picks the instantiation for (tier, deposit, probe, resident CTAs per SM) and applies `fn` to it.
template <typename Fn>
cudaError_t with_trace_kernel(int tier, int deposit, bool probe, int min_blocks, bool count, Fn fn, int room_steps = 2)
{
    if (count && !probe) {      // counting variant: one instantiation per tier
        if (tier == kTierRooms) return fn(k_trace<kTierRooms, FMGI_DEPOSIT_VEC4, false, 4, true>);
        if (tier == FMGI_TIER_GRID) return fn(k_trace<FMGI_TIER_GRID, FMGI_DEPOSIT_VEC4, false, 4, true>);
        if (tier == kTierSoupPlanes) return fn(k_trace<kTierSoupPlanes, FMGI_DEPOSIT_VEC4, false, 4, true>);
    }
#define FMGI_PICK_DEPOSIT(T, B)                                                                  \
    switch (deposit) {                                                                           \
        case FMGI_DEPOSIT_SCALAR: return fn(k_trace<T, FMGI_DEPOSIT_SCALAR, false, B>);          \
        case FMGI_DEPOSIT_WARP_AGG: return fn(k_trace<T, FMGI_DEPOSIT_WARP_AGG, false, B>);      \
        default: return fn(k_trace<T, FMGI_DEPOSIT_VEC4, false, B>);                             \
    }
    if (tier == kTierRooms) {
        if (probe) return fn(k_trace<kTierRooms, FMGI_DEPOSIT_VEC4, true, 3>);
        if (deposit == FMGI_DEPOSIT_VEC4) {
            // boxes per iteration of the photon loop: 3 while the box table is L1-resident (example.png: 6.16 vs 6.27 ms),
            // 2 for big scenes, where every further box is an L2 round trip the lanes that are ready to shade wait for
            // (synth4000: 76.4 vs 78.7 ms); FMGI_ROOM_STEPS = 1..4 or 64 (whole walk) overrides
            static const int forced = getenv("FMGI_ROOM_STEPS") ? atoi(getenv("FMGI_ROOM_STEPS")) : 0;
            const int steps = forced ? forced : room_steps;
            if (steps == 1) return fn(k_trace<kTierRooms, FMGI_DEPOSIT_VEC4, false, 4, false, 1>);
            if (steps == 3) return fn(k_trace<kTierRooms, FMGI_DEPOSIT_VEC4, false, 4, false, 3>);
            if (steps == 4) return fn(k_trace<kTierRooms, FMGI_DEPOSIT_VEC4, false, 4, false, 4>);
            if (steps >= 64) return fn(k_trace<kTierRooms, FMGI_DEPOSIT_VEC4, false, 4, false, kRoomMaxSteps>);
        }
        FMGI_PICK_DEPOSIT(kTierRooms, 4)
    }
    if (tier == FMGI_TIER_GRID) {
        if (probe) return fn(k_trace<FMGI_TIER_GRID, FMGI_DEPOSIT_VEC4, true, 3>);
        if (min_blocks == 3) { FMGI_PICK_DEPOSIT(FMGI_TIER_GRID, 3) }
        FMGI_PICK_DEPOSIT(FMGI_TIER_GRID, 4)
    }
    if (tier == kTierSoupPlanes) {
        if (probe) return fn(k_trace<kTierSoupPlanes, FMGI_DEPOSIT_VEC4, true, 3>);
        if (min_blocks == 3) { FMGI_PICK_DEPOSIT(kTierSoupPlanes, 3) }
        FMGI_PICK_DEPOSIT(kTierSoupPlanes, 4)
    }
    if (probe) return fn(k_trace<FMGI_TIER_SOUP, FMGI_DEPOSIT_VEC4, true, 3>);
    if (min_blocks == 3) { FMGI_PICK_DEPOSIT(FMGI_TIER_SOUP, 3) }
    FMGI_PICK_DEPOSIT(FMGI_TIER_SOUP, 4)
#undef FMGI_PICK_DEPOSIT
}

// The pooled kernel (trace_pool.cuh) for K rays per lane.
template <typename Fn>
cudaError_t with_pool_kernel(int k, Fn fn)
{
    switch (k) {
        case 2: return fn(k_trace_pool<2, FMGI_DEPOSIT_VEC4, 8>, 2);
        case 3: return fn(k_trace_pool<3, FMGI_DEPOSIT_VEC4, 7>, 3);
        default: return fn(k_trace_pool<4, FMGI_DEPOSIT_VEC4, 5>, 4);
    }
}

cudaError_t launch_trace(fmgi_scene *s, const TraceParams &p, int deposit, bool probe, int blocks, cudaStream_t st,
                         bool count = false)
{
    return with_trace_kernel(s->kernel_tier, deposit, probe, s->min_blocks, count, [&](auto kernel) {
        cudaError_t e = cudaSuccess;
        if (s->smem_bytes > 48 * 1024)
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes);
        if (e != cudaSuccess) return e;
        kernel<<<blocks, kTraceThreads, s->smem_bytes, st>>>(p);
        s->launches++;
        return cudaGetLastError();
    }, s->build->rooms.boxes.size() <= 1024 ? 3 : 2);
}


// device attributes do not change; cudaDevAttrClockRate in particular is slow to query (and
// cudaGetDeviceProperties costs about a millisecond per call)
struct DevAttr { int valid, sms, smem_optin, clock_khz; };
int device_attrs(int device, DevAttr &out)
{
    static DevAttr cache[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    DevAttr &a = cache[device & 63];
    if (!a.valid) {
        FMGI_CUDA(cudaDeviceGetAttribute(&a.sms, cudaDevAttrMultiProcessorCount, device));
        FMGI_CUDA(cudaDeviceGetAttribute(&a.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        FMGI_CUDA(cudaDeviceGetAttribute(&a.clock_khz, cudaDevAttrClockRate, device));
        a.valid = 1;
    }
    out = a;
    return FMGI_OK;
}

// Host half of fmgi_scene_create: rectangle tables, tier choice, floor-plan grid.  No device work, so one
// build serves every GPU of a multi-GPU bake.
int build_host(std::shared_ptr<HostBuild> &out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows,
               int num_windows, const fmgi_rect *lights, int num_lights, int num_texels, const fmgi_options &o)
{
    if ((num_walls && !walls) || (num_windows && !windows) || (num_lights && !lights))
        return fail(FMGI_ERR_ARG, "NULL rectangle table with non-zero count");
    int ndev = 0;
    FMGI_CUDA(cudaGetDeviceCount(&ndev));
    if (o.device < 0 || o.device >= ndev) return fail(FMGI_ERR_ARG, "device ordinal out of range");
    DevAttr attr;
    if (int rc = device_attrs(o.device, attr)) return rc;

    auto b = std::make_shared<HostBuild>();
    int tier = o.tier;
    if (const char *v = getenv("FMGI_TIER")) tier = atoi(v);
    const bool want_rooms = tier == FMGI_TIER_ROOMS;
    const bool auto_tier = tier != FMGI_TIER_SOUP && tier != FMGI_TIER_GRID && tier != FMGI_TIER_ROOMS;
    // A big scene that will (try to) use the room tier: its box decomposition (rooms_build.cpp, 12 ms for 21.5k
    // rectangles) does not depend on the rectangle tables (2 ms) - both at once.
    const char *refused = nullptr;
    std::thread rooms_thread;
    if ((want_rooms || auto_tier) && num_walls >= 2048)
        rooms_thread = std::thread([&]() {
            const double tr0 = now_ms();
            refused = build_rooms(b->rooms, walls, num_walls, windows, num_windows, lights, num_lights);
            b->rooms_ms = now_ms() - tr0;
        });
    const double t0 = now_ms();
    const char *why = prepare_scene(b->scene, walls, num_walls, windows, num_windows, lights, num_lights, num_texels);
    if (why[0]) {
        if (rooms_thread.joinable()) rooms_thread.join();
        return fail(FMGI_ERR_ARG, why);
    }
    b->wall_area.resize(num_walls);
    b->wall_floor.resize(num_walls);
    b->wall_wh.reserve((size_t)6 * num_walls);
    for (int i = 0; i < num_walls; i++) {
        const ShadeRect &sh = b->scene.shade[i];
        b->wall_area[i] = sh.wlen * sh.hlen;                                            // getArea, rectangle.c:194-197
        b->wall_floor[i] = walls[i].pos[2] == 0 && walls[i].width[2] == 0 && walls[i].height[2] == 0;
        for (int c = 0; c < 3; c++) b->wall_wh.push_back(walls[i].width[c]);
        for (int c = 0; c < 3; c++) b->wall_wh.push_back(walls[i].height[c]);
    }
    b->prepare_ms = now_ms() - t0;
    if (rooms_thread.joinable()) rooms_thread.join();

    // tier: brute force over the shared-memory soup for small scenes, floor-plan grid otherwise
    HostScene &hs = b->scene;
    const size_t soup_bytes = hs.axis.size() * sizeof(AxisPairBlock) + hs.general.size() * sizeof(GeneralRect);
    const int colliders = hs.num_axis_rects + (int)hs.general.size();
    // AUTO: the brute-force soup only for a handful of colliders (one bare room); everything else walks the box
    // decomposition of the room tier when every collider is axis parallel (all parseLayout output is), else the
    // floor-plan grid (measured on the 172-rectangle example.png scene the grid is 20 % faster than the soup + plane
    // tables, and the room tier 1.8x faster than the grid)
    if (auto_tier)
        tier = (colliders <= 64 && soup_bytes <= (size_t)attr.smem_optin) ? FMGI_TIER_SOUP : FMGI_TIER_ROOMS;
    if (tier == FMGI_TIER_ROOMS) {
        if (!refused) {
            const double tr0 = now_ms();
            refused = build_rooms(b->rooms, walls, num_walls, windows, num_windows, lights, num_lights);
            b->rooms_ms = now_ms() - tr0;
        }
        if (refused[0]) {
            if (want_rooms) return fail(FMGI_ERR_UNSUPPORTED, std::string("room tier: ") + refused);
            b->rooms = RoomScene();
            tier = FMGI_TIER_GRID;
        }
    } else {
        b->rooms = RoomScene();
    }
    b->tier = tier;
    b->kernel_tier = tier;
    float cell = 0.0f;
    if (const char *v = getenv("FMGI_GRID_CELL")) cell = (float)atof(v);
    // The per-collider half of the grid build always runs here; the per-cell half (binning, ordering, laying out T)
    // runs on the device for scenes of a few thousand colliders and more (grid_build.cuh) - 11 ms on the host for
    // 21.5k rectangles, a fraction of a millisecond on the GPU.  FMGI_GRID_BUILD=host|device overrides.
    const double t1 = now_ms();
    if (tier != FMGI_TIER_ROOMS) {
        grid_classify(hs, walls, num_walls, windows, num_windows, lights, num_lights, cell, b->grid_items);
        b->device_grid = num_walls >= 2048;
        if (const char *v = getenv("FMGI_GRID_BUILD")) b->device_grid = v[0] == 'd';
        if (!b->device_grid) {
            grid_assemble_host(hs, b->grid_items);
            b->grid_items.clear();
            b->grid_items.shrink_to_fit();
        }
    }
    b->grid_ms = now_ms() - t1 + b->rooms_ms;
    if (tier == FMGI_TIER_SOUP) {
        b->smem_bytes = soup_bytes;
        // each lane walks one of the two blocks of every pair: two rectangle tests per pair
        b->tests_per_ray = hs.axis.size() + hs.general.size();
        // horizontal rectangles through the plane tables when all of them fit (and there are any)
        bool planes = hs.grid_overflow_horizontal == 0 && hs.pair_begin[3] > hs.pair_begin[2];
        if (const char *v = getenv("FMGI_SOUP_PLANES")) planes = planes && atoi(v) != 0;       // tuning knob
        if (planes) {
            b->kernel_tier = kTierSoupPlanes;
            b->tests_per_ray = 2 * (size_t)hs.pair_begin[2] + hs.general.size();   // x and y lists only
        }
    }
    out = b;
    return FMGI_OK;
}

// Device half: uploads the tables of a host build to o.device.
int scene_from_build(fmgi_scene **out, std::shared_ptr<HostBuild> b, const fmgi_options &o)
{
    *out = nullptr;
    DevAttr attr;
    if (int rc = device_attrs(o.device, attr)) return rc;
    // a scene that fails half-way gives its device blocks back through fmgi_scene_destroy
    std::unique_ptr<fmgi_scene, void (*)(fmgi_scene *)> s(new fmgi_scene(b), fmgi_scene_destroy);
    s->device = o.device;
    s->num_sms = attr.sms; s->clock_khz = attr.clock_khz;
    s->tier = b->tier; s->kernel_tier = b->kernel_tier;
    s->smem_bytes = b->smem_bytes; s->tests_per_ray = b->tests_per_ray;
    DeviceGuard guard(o.device);
    if (s->kernel_tier == kTierRooms) {
        FMGI_CUDA(upload(&s->d_room_boxes, b->rooms.boxes));
        FMGI_CUDA(upload(&s->d_room_face_grids, b->rooms.face_grids));
        FMGI_CUDA(upload(&s->d_room_face_cells, b->rooms.face_cells));
        FMGI_CUDA(upload(&s->d_room_bounds, b->rooms.bounds));
        FMGI_CUDA(upload(&s->d_room_nodes, b->rooms.nodes));
        FMGI_CUDA(upload(&s->d_room_starts, b->rooms.starts));
    } else if (s->kernel_tier != FMGI_TIER_SOUP) {
        if (b->device_grid) {
            const double tg0 = now_ms();
            FMGI_CUDA(grid_assemble_device(s->host.grid, b->grid_items, &s->d_grid_table, &s->grid_records, nullptr));
            s->grid_device_ms = now_ms() - tg0;
        } else {
            FMGI_CUDA(upload(&s->d_grid_table, s->host.grid_table));
            s->grid_records = s->host.grid_table.size();
        }
    }
    FMGI_CUDA(upload(&s->d_axis, s->host.axis));
    FMGI_CUDA(upload(&s->d_general, s->host.general));
    FMGI_CUDA(upload(&s->d_shade, s->host.shade));
    FMGI_CUDA(upload(&s->d_emitters, s->host.emitters));
    const size_t E = s->host.emitters.size();
    MemPool &pool = MemPool::get();
    FMGI_CUDA(pool.alloc((void **)&s->d_jobs, job_table_words(E) * sizeof(unsigned long long), false));
    s->job_tables = 1;
    FMGI_CUDA(pool.alloc((void **)&s->d_counters, 8 * sizeof(unsigned long long), false));
    FMGI_CUDA(pool.alloc((void **)&s->h_jobs, job_table_words(E) * sizeof(unsigned long long), true));
    FMGI_CUDA(pool.alloc((void **)&s->h_counters, 8 * sizeof(unsigned long long), true));
    memset(s->h_counters, 0, 8 * sizeof(unsigned long long));
    FMGI_CUDA(cudaEventCreate(&s->ev_start));
    FMGI_CUDA(cudaEventCreate(&s->ev_stop));

    if (const char *v = getenv("FMGI_TUNE_BLOCKS")) s->min_blocks = atoi(v) == 3 ? 3 : 4;   // tuning knob
    fmgi_scene *sp = s.get();
    FMGI_CUDA(with_trace_kernel(s->kernel_tier, FMGI_DEPOSIT_VEC4, false, s->min_blocks, false, [&](auto kernel) {
        cudaError_t e = cudaSuccess;
        if (sp->smem_bytes > 48 * 1024)
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp->smem_bytes);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sp->blocks_per_sm, kernel, kTraceThreads, sp->smem_bytes);
        return e;
    }));
    if (s->blocks_per_sm < 1) return fail(FMGI_ERR_CUDA, "trace kernel does not fit on an SM");
    // pooled kernel: grid tier without misc records (those leave the PTX walk loop for a slow path)
    // opt-in (FMGI_POOL_K = 2..4 rays per lane): measured on example.png the pool needs 0.28 warp iterations of the walk
    // loop per ray instead of 0.52, but its switch code and the shared memory it takes from L1 still cost more
    int pool_k = 0;
    if (const char *v = getenv("FMGI_POOL_K")) pool_k = atoi(v);
    if (s->kernel_tier == FMGI_TIER_GRID && s->host.grid_misc == 0 && pool_k >= 2 &&
        (int)s->host.emitters.size() < kPoolMaxEmitters) {
        pool_k = pool_k > 4 ? 4 : pool_k;
        FMGI_CUDA(with_pool_kernel(pool_k, [&](auto kernel, int k) {
            const size_t smem = (size_t)(kPoolThreads / 32) * 32 * k * kPoolSlotBytes;
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            // shared-memory carve-out: just what the resident CTAs need (FMGI_POOL_CARVEOUT = percent overrides) - the
            // rest of the 256 KB stays L1 for the grid table
            int carve = -1;
            if (const char *v = getenv("FMGI_POOL_CARVEOUT")) carve = atoi(v);
            if (e == cudaSuccess && carve >= 0)
                e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            if (e == cudaSuccess)
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sp->pool_blocks_per_sm, kernel, kPoolThreads, smem);
            return e;
        }));
        if (s->pool_blocks_per_sm >= 1) s->pool_k = pool_k;
    }
    *out = s.release();
    return FMGI_OK;
}

}  // namespace

extern "C" {

void fmgi_default_options(fmgi_options *opt)
{
    if (!opt) return;
    memset(opt, 0, sizeof *opt);
    opt->struct_size = sizeof *opt;
    opt->max_depth = 8;          // photonmap.c:173
    opt->seed = 1;
    opt->num_gpus = 1;
    opt->num_shards = 1;
    opt->tier = FMGI_TIER_AUTO;
    opt->deposit = FMGI_DEPOSIT_VEC4;
}

const char *fmgi_last_error(void) { return g_last_error.c_str(); }
#ifndef FMGI_SRC_HASH
#define FMGI_SRC_HASH "unknown"
#endif
const char *fmgi_version(void) { return "fmgi-b200 0.2 (sm_100a) src " FMGI_SRC_HASH; }
const char *fmgi_source_hash(void) { return FMGI_SRC_HASH; }

void fmgi_release_cache(void) { MemPool::get().release(); }
uint64_t fmgi_cached_bytes(void) { return MemPool::get().cached_bytes(); }

int fmgi_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fmgi_scene_create(fmgi_scene **out, const fmgi_rect *walls, int num_walls, const fmgi_rect *windows,
                      int num_windows, const fmgi_rect *lights, int num_lights, int num_texels,
                      const fmgi_options *opt)
{
    if (!out) return fail(FMGI_ERR_ARG, "out is NULL");
    *out = nullptr;
    const fmgi_options o = resolve(opt);
    std::shared_ptr<HostBuild> b;
    const int rc = build_host(b, walls, num_walls, windows, num_windows, lights, num_lights, num_texels, o);
    if (rc) return rc;
    return scene_from_build(out, b, o);
}

void fmgi_scene_destroy(fmgi_scene *s)
{
    if (!s) return;
    DeviceGuard guard(s->device);
    if (s->traced) cudaEventSynchronize(s->ev_stop);      // blocks go back to the pool: nothing may still use them
    for (cudaStream_t st : s->streams) cudaStreamSynchronize(st);
    MemPool &pool = MemPool::get();
    pool.free(s->d_axis); pool.free(s->d_general); pool.free(s->d_shade); pool.free(s->d_emitters);
    pool.free(s->d_grid_table);
    pool.free(s->d_room_boxes); pool.free(s->d_room_face_grids); pool.free(s->d_room_face_cells); pool.free(s->d_room_bounds); pool.free(s->d_room_nodes);
    pool.free(s->d_room_starts);
    pool.free(s->d_jobs); pool.free(s->d_counters); pool.free(s->d_scratch);
    pool.free(s->d_tile_walls); pool.free(s->h_tile_walls);
    pool.free(s->d_png_walls); pool.free(s->h_png_walls);
    pool.free(s->d_ao); pool.free(s->d_ao_walls);
    pool.free(s->h_jobs); pool.free(s->h_counters);
    if (s->ev_start) cudaEventDestroy(s->ev_start);
    if (s->ev_stop) cudaEventDestroy(s->ev_stop);
    delete s;
}

uint64_t fmgi_scene_photon_count(const fmgi_scene *s, int spa, const fmgi_options *opt)
{
    if (!s) return 0;
    const fmgi_options o = resolve(opt);
    std::vector<unsigned long long> jobs(job_table_words(s->host.emitters.size()));
    return fill_jobs(s, spa, o, jobs.data());
}

int fmgi_scene_trace(fmgi_scene *s, void *atlas_dev, int spa, const fmgi_options *opt, void *cuda_stream)
{
    if (!s || !atlas_dev) return fail(FMGI_ERR_ARG, "scene or atlas is NULL");
    const fmgi_options o = resolve(opt);
    if (o.shard < 0 || o.shard >= o.num_shards) return fail(FMGI_ERR_ARG, "shard out of range");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int E = (int)s->host.emitters.size();
    const size_t table_words = job_table_words((size_t)E);

    if (s->traced) FMGI_CUDA(cudaEventSynchronize(s->ev_stop));   // pinned staging is reused

    // fp32 accumulation passes.  Every texel is a float sum of deposits of about 16; once a hot texel
    // passes ~1e7 a single deposit is a few ulps and the sum drifts (measured: -0.054 % total energy at
    // 4e9 photons on example.png; the reference has the same flaw, photonmap.c:251, at 1e3 x fewer
    // photons per hour).  Above kAccumPhotons the shard is therefore traced in several passes into a
    // zeroed scratch atlas that is added to the caller's atlas after each pass.  A pass is just a finer
    // shard of the photon index space, so the sample set does not change.
    std::vector<unsigned long long> whole(table_words);
    const unsigned long long total_all = fill_jobs(s, spa, o, whole.data());
    // limits of the packed Philox counter word (philox.cuh)
    if (o.max_depth > kPhiloxMaxDepth) return fail(FMGI_ERR_UNSUPPORTED, "more than 15 bounces per photon");
    if (E >= kPhiloxMaxEmitters) return fail(FMGI_ERR_UNSUPPORTED, "more than 2^20 emitters");
    for (int e = 0; e < E; e++)
        if (whole[E + 1 + e] + whole[2 * E + 1 + e] > kPhiloxMaxPhotons)
            return fail(FMGI_ERR_UNSUPPORTED, "more than 2^40 photons from one emitter");
    int passes = (int)((total_all + kAccumPhotons - 1) / kAccumPhotons);
    if (const char *v = getenv("FMGI_ACCUM_PASSES")) passes = atoi(v);     // tuning / test knob
    if (passes < 1) passes = 1;
    bool count_tests = o.count_tests != 0;
    if (const char *v = getenv("FMGI_COUNT_TESTS")) count_tests = atoi(v) != 0;
    MemPool &pool = MemPool::get();
    if ((size_t)passes > s->job_tables) {
        pool.free(s->d_jobs); pool.free(s->h_jobs);
        s->d_jobs = nullptr; s->h_jobs = nullptr;
        FMGI_CUDA(pool.alloc((void **)&s->d_jobs, passes * table_words * sizeof(unsigned long long), false));
        FMGI_CUDA(pool.alloc((void **)&s->h_jobs, passes * table_words * sizeof(unsigned long long), true));
        s->job_tables = (size_t)passes;
    }
    const size_t atlas_bytes = (size_t)s->host.num_texels * sizeof(float4);
    if (passes > 1 && !s->d_scratch)
        FMGI_CUDA(pool.alloc((void **)&s->d_scratch, atlas_bytes ? atlas_bytes : 16, false));

    // Chunk size: kChunkPhotons for big bakes; a smaller bake is cut finer, down to 32 photons, so that every
    // resident warp gets at least 32 chunks: the last chunk of a warp is a serial tail (256 photons on 32 lanes are
    // 8 photons x 5 rays per lane) that the other warps cannot help with.  Measured on example.png, 8 bounces:
    // 1.25e7 photons 2.80 ms with 256-photon chunks, 2.71 ms with 64; 1.25e6 photons 0.44 vs 0.33 ms (32); at 1e8
    // photons the big chunk wins by 0.7 % (fewer claims).  Round 1 used 256 always and sized the grid by the
    // chunk count: a 1.25e5-photon shard then ran 61 CTAs with one serial chunk per warp (0.23 ms, now 0.07).
    const unsigned long long wave = (unsigned long long)s->num_sms * s->blocks_per_sm;
    const unsigned long long per_pass = total_all / (unsigned long long)passes;
    int chunk = kChunkPhotons;
    while (chunk > kMinChunkPhotons && per_pass / chunk < 32 * wave * (kTraceThreads / 32)) chunk >>= 1;
    if (const char *v = getenv("FMGI_CHUNK")) { const int c = atoi(v); if (c >= 1 && c <= 65536) chunk = c; }   // tuning knob
    fmgi_options op = o;
    op.num_shards = o.num_shards * passes;
    std::vector<unsigned long long> totals(passes), chunks(passes);
    for (int c = 0; c < passes; c++) {
        op.shard = o.shard * passes + c;
        totals[c] = fill_jobs(s, spa, op, s->h_jobs + c * table_words, &chunks[c], chunk);
    }
    FMGI_CUDA(cudaMemcpyAsync(s->d_jobs, s->h_jobs, passes * table_words * sizeof(unsigned long long),
                              cudaMemcpyHostToDevice, st));
    FMGI_CUDA(cudaMemsetAsync(s->d_counters, 0, 8 * sizeof(unsigned long long), st));

    FMGI_CUDA(cudaEventRecord(s->ev_start, st));
    for (int c = 0; c < passes; c++) {
        if (totals[c] == 0) continue;
        TraceParams p = base_params(s);
        p.job_begin = s->d_jobs + c * table_words;
        p.photon_first = p.job_begin + (E + 1);
        p.photon_count = p.job_begin + (2 * E + 1);
        p.total_jobs = chunks[c];
        p.chunk = chunk;
        p.atlas = reinterpret_cast<float4 *>(passes > 1 ? (void *)s->d_scratch : atlas_dev);
        p.max_depth = o.max_depth;
        set_seed(p, o.seed);
        if (passes > 1) {
            FMGI_CUDA(cudaMemsetAsync(s->d_scratch, 0, atlas_bytes, st));
            FMGI_CUDA(cudaMemsetAsync(s->d_counters + 4, 0, sizeof(unsigned long long), st));   // work counter
        }
        // Pooled kernel (trace_pool.cuh) when the pass keeps every resident warp's pool busy for a while and the
        // photon ids fit its packed word; k_trace otherwise (small bakes, soup tiers, misc records, counting).
        const unsigned long long pool_warps = (unsigned long long)s->num_sms * s->pool_blocks_per_sm * (kPoolThreads / 32);
        bool pooled = s->pool_k > 0 && !count_tests && o.deposit == FMGI_DEPOSIT_VEC4 && o.max_depth <= kPoolMaxDepth &&
                      totals[c] >= 8ull * pool_warps * 32 * s->pool_k;
        for (int e = 0; pooled && e < E; e++) {
            const unsigned long long *job = s->h_jobs + c * table_words;
            pooled = job[E + 1 + e] + job[2 * E + 1 + e] < kPoolMaxPhotonIndex;
        }
        if (const char *v = getenv("FMGI_POOL")) pooled = pooled && atoi(v) != 0;       // tuning / test knob
        if (pooled) {
            FMGI_CUDA(with_pool_kernel(s->pool_k, [&](auto kernel, int k) {
                const size_t smem = (size_t)(kPoolThreads / 32) * 32 * k * kPoolSlotBytes;
                const int blocks = s->num_sms * s->pool_blocks_per_sm;
                kernel<<<blocks, kPoolThreads, smem, st>>>(p);
                s->launches++;
                return cudaGetLastError();
            }));
        } else {
            // persistent grid: one wave of resident CTAs, never more warps than chunks of work
            unsigned long long want = (chunks[c] * 32 + kTraceThreads - 1) / kTraceThreads;
            const int blocks = (int)(want < wave ? (want ? want : 1) : wave);
            FMGI_CUDA(launch_trace(s, p, o.deposit, false, blocks, st, count_tests));
        }
        s->pooled_last = pooled;
        if (passes > 1 && s->host.num_texels > 0) {
            k_accumulate<<<s->num_sms * 8, 256, 0, st>>>(reinterpret_cast<float4 *>(atlas_dev), s->d_scratch,
                                                         (size_t)s->host.num_texels);
            s->launches++;
            FMGI_CUDA(cudaGetLastError());
        }
    }
    FMGI_CUDA(cudaEventRecord(s->ev_stop, st));
    FMGI_CUDA(cudaMemcpyAsync(s->h_counters, s->d_counters, 8 * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
    s->note_stream(st);
    s->traced = true;
    return FMGI_OK;
}

int fmgi_scene_sync(fmgi_scene *s, fmgi_stats *stats)
{
    if (!s) return fail(FMGI_ERR_ARG, "scene is NULL");
    DeviceGuard guard(s->device);
    FMGI_CUDA(cudaStreamSynchronize(s->last_stream));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->photons = s->h_counters[0];
        stats->rays = s->h_counters[1];
        stats->deposits = s->h_counters[2];
        stats->mirror_bounces = s->h_counters[3];
        stats->rect_tests = s->h_counters[1] * s->tests_per_ray + s->h_counters[5];   // static scan + counted lookups
        stats->kernel_launches = s->launches;
        if (s->traced) {
            float ms = 0;
            FMGI_CUDA(cudaEventElapsedTime(&ms, s->ev_start, s->ev_stop));
            stats->trace_ms = ms;
        }
        stats->num_gpus = 1;
        stats->prepare_ms = s->build->prepare_ms;
        stats->grid_build_ms = s->build->grid_ms + s->grid_device_ms;
        stats->pool_rays = s->pooled_last ? s->pool_k : 0;
#ifdef FMGI_CHECKED
        stats->bounds_violations = (int32_t)std::min<unsigned long long>(s->h_counters[6], 0x7fffffffull);
        if (s->h_counters[6])
            fprintf(stderr, "[fmgi] bounds-checked build: %llu index violations, first at site %llu\n", s->h_counters[6],
                    s->h_counters[7]);
#else
        stats->bounds_violations = -1;
#endif
        stats->tier = s->tier;
        stats->num_sms = s->num_sms;
        stats->sm_clock_khz = s->clock_khz;
    }
    return FMGI_OK;
}

// ---- host-buffer bake ---------------------------------------------------------------------------------

} // extern "C"

namespace {

int tonemap_impl(fmgi_scene *s, const void *atlas_dev, int spa, int tint_extra, void *out_dev, cudaStream_t st, bool png);

// ---- process-wide per-GPU runtime: streams, events and peer mappings are created once ---------------------
//
// performGlobalIlluminationCl is a one-shot call, but a process that bakes more than once (a harness, a
// service) must not pay stream / event creation and cudaDeviceEnablePeerAccess on every call; the primary
// contexts themselves persist for the life of the process anyway.
struct GpuRuntime {
    struct Dev {
        bool ready = false;
        cudaStream_t trace = nullptr, copy = nullptr;
        cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // upload 0/1, fold 2/3, read-back 4/5
    };
    std::mutex mu;
    Dev dev[64];
    signed char peer[64][64];      // 0 unknown, 1 enabled (row device maps column device's memory), -1 unavailable
    GpuRuntime() { memset(peer, 0, sizeof peer); }
    static GpuRuntime &get() { static GpuRuntime r; return r; }

    // current device must be d
    cudaError_t ensure(int d)
    {
        Dev &v = dev[d & 63];
        if (v.ready) return cudaSuccess;
        cudaError_t e = cudaStreamCreateWithFlags(&v.trace, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v.copy, cudaStreamNonBlocking);
        for (int i = 0; i < 6 && e == cudaSuccess; i++) e = cudaEventCreate(&v.ev[i]);
        if (e == cudaSuccess) v.ready = true;
        return e;
    }
    // current device must be a; true when a can read b's memory directly
    bool map_peer(int a, int b)
    {
        if (a == b) return true;
        std::lock_guard<std::mutex> lock(mu);
        signed char &st = peer[a & 63][b & 63];
        if (st == 0) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (can) {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(b, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) can = 0;
                cudaGetLastError();
            }
            st = can ? 1 : -1;
        }
        return st == 1;
    }
};

template <typename Fn>
void parallel_for(int n, Fn fn)
{
    if (n == 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (int g = 0; g < n; g++) th.emplace_back(fn, g);
    for (auto &t : th) t.join();
}

// Host-buffer bake.  tiles_out == NULL: geo->texels receives the float atlas (fmgi_bake); otherwise the
// atlas is tone-mapped on GPU 0 and only the packed RGB tiles come back (fmgi_bake_tiles).
//
// Shape (G GPUs, one host thread each):
//   1. every GPU traces its photon shard into its own ZEROED device atlas; at the same time GPU g uploads
//      slice g of the caller's atlas (texels [lo_g, hi_g)) on a second stream, so the upload hides behind
//      the trace (global_illumination_cl.c:295 initialises the device buffer from the host array);
//   2. reduce-scatter over peer memory: GPU g sums slice g of ALL atlases, reading the other GPUs' copies
//      through NVLink peer mappings, plus the uploaded host slice (k_fold_slice) - every GPU pulls
//      (G-1)/G of one atlas at the same time instead of GPU 0 pulling G-1 atlases;
//   3. every GPU writes its slice straight back into the caller's atlas over its own PCIe link.
int bake_impl(struct Geometry *geo_, int spa, const fmgi_options *opt, fmgi_stats *stats, uint8_t *tiles_out,
              int tint_extra, bool one_shot = false, bool tiles_png = false)
{
    fmgi_geometry *geo = reinterpret_cast<fmgi_geometry *>(geo_);
    if (!geo) return fail(FMGI_ERR_ARG, "geo is NULL");
    if (geo->numTexels < 0 || (geo->numTexels && !geo->texels)) return fail(FMGI_ERR_ARG, "bad texel atlas");
    const double t_begin = now_ms();
    fmgi_options o = resolve(opt);
    int ndev = 0;
    FMGI_CUDA(cudaGetDeviceCount(&ndev));               // the first CUDA call of a process initialises the driver
    const double driver_ms = now_ms() - t_begin;
    if (ndev < 1) return fail(FMGI_ERR_CUDA, "no CUDA device");
    if (o.device < 0 || o.device >= ndev) return fail(FMGI_ERR_ARG, "device ordinal out of range");
    const int G = std::min(std::min(o.num_gpus, ndev), kMaxFoldPeers + 1);
    const size_t num_texels = (size_t)geo->numTexels;
    // the caller's current device is restored on every exit path (the workers and the fold switch devices)
    const double tg0 = now_ms();
    DeviceGuard entry_guard(o.device);
    cudaFree(nullptr);                                  // primary context of the first GPU (first call of a process)
    const double context_ms = now_ms() - tg0;

    std::shared_ptr<HostBuild> build;
    const double tb0 = now_ms();
    if (int rc = build_host(build, geo->walls, geo->numWalls, geo->windows, geo->numWindows, geo->lights, geo->numLights,
                            geo->numTexels, o))
        return rc;
    const double build_ms = now_ms() - tb0;

    struct PerGpu {
        int device = 0;
        fmgi_scene *scene = nullptr;
        float4 *atlas = nullptr;      // this GPU's deposits (whole atlas)
        float4 *init = nullptr;       // the caller's values of this GPU's slice
        float4 *staged = nullptr;     // peers' slices copied over when no peer mapping exists
        size_t lo = 0, hi = 0;        // slice [lo, hi) in texels
        fmgi_stats st = {};
        int rc = FMGI_OK;
        std::string err;
        double init_ms = 0, create_ms = 0, sync_ms = 0, alloc_ms = 0, h2d_call_ms = 0, trace_call_ms = 0;
        float h2d_ms = 0, fold_ms = 0, d2h_ms = 0;
    };
    std::vector<PerGpu> gpus(G);
    GpuRuntime &rt = GpuRuntime::get();
    for (int g = 0; g < G; g++) {
        gpus[g].device = (o.device + g) % ndev;          // G <= ndev: distinct GPUs, starting at the caller's
        // slices in units of 64 texels (1 KiB)
        const size_t units = (num_texels + 63) / 64;
        gpus[g].lo = std::min(num_texels, units * g / G * 64);
        gpus[g].hi = std::min(num_texels, units * (g + 1) / G * 64);
    }

    // ---- phase 1: upload + trace ---------------------------------------------------------------------------
    parallel_for(G, [&](int g) {
        PerGpu &me = gpus[g];
        auto bail = [&](int rc) { me.rc = rc; me.err = g_last_error; };
        fmgi_options og = o;
        og.device = me.device;
        // the caller's shard is subdivided over this call's GPUs
        og.num_shards = o.num_shards * G;
        og.shard = o.shard * G + g;
        const double ti0 = now_ms();
        if (cudaSetDevice(me.device) != cudaSuccess || rt.ensure(me.device) != cudaSuccess)
            return bail(fail(FMGI_ERR_CUDA, "cudaSetDevice / stream creation failed"));
        me.init_ms = now_ms() - ti0;                      // context creation on the first call of the process
        GpuRuntime::Dev &dv = rt.dev[me.device & 63];
        const double tc0 = now_ms();
        int rc = scene_from_build(&me.scene, build, og);
        if (rc) return bail(rc);
        me.create_ms = now_ms() - tc0;
        cudaSetDevice(me.device);
        const size_t slice = me.hi - me.lo;
        MemPool &pool = MemPool::get();
        const double ta0 = now_ms();
        if (pool.alloc((void **)&me.atlas, std::max<size_t>(num_texels, 1) * sizeof(float4), false) != cudaSuccess ||
            pool.alloc((void **)&me.init, std::max<size_t>(slice, 1) * sizeof(float4), false) != cudaSuccess)
            return bail(fail(FMGI_ERR_CUDA, "atlas allocation failed"));
        me.alloc_ms = now_ms() - ta0;
        cudaError_t e = cudaMemsetAsync(me.atlas, 0, num_texels * sizeof(float4), dv.trace);
        if (e != cudaSuccess) return bail(fail(FMGI_ERR_CUDA, std::string("atlas clear: ") + cudaGetErrorString(e)));
        // the trace is enqueued FIRST: an upload from pageable memory (what main.c hands us, parseLayout.c:526)
        // blocks the host thread while the driver stages it, and then runs under the kernel instead of before it
        const double ts0 = now_ms();
        rc = fmgi_scene_trace(me.scene, me.atlas, spa, &og, dv.trace);
        if (rc) return bail(rc);
        me.trace_call_ms = now_ms() - ts0;
        cudaSetDevice(me.device);
        // the caller's atlas (CL_MEM_COPY_HOST_PTR, global_illumination_cl.c:295) rides on the copy stream - but not
        // before the trace kernel has started: the trace stream's own small upload (the job tables) sits behind the
        // atlas clear, and an atlas upload that reaches the copy engine first holds it up for its whole length
        // (measured on the 1.83 GB atlas: kernel start 33 ms late)
        e = cudaStreamWaitEvent(dv.copy, me.scene->ev_start, 0);
        if (e == cudaSuccess) e = cudaEventRecord(dv.ev[0], dv.copy);
        const double th0 = now_ms();
        if (e == cudaSuccess && slice)
            e = cudaMemcpyAsync(me.init, geo->texels + 4 * me.lo, slice * sizeof(float4), cudaMemcpyHostToDevice, dv.copy);
        me.h2d_call_ms = now_ms() - th0;
        if (e == cudaSuccess) e = cudaEventRecord(dv.ev[1], dv.copy);
        if (e != cudaSuccess) return bail(fail(FMGI_ERR_CUDA, std::string("atlas upload: ") + cudaGetErrorString(e)));
        rc = fmgi_scene_sync(me.scene, &me.st);
        if (rc) return bail(rc);
        me.sync_ms = now_ms() - ts0;
        cudaSetDevice(me.device);
        e = cudaStreamSynchronize(dv.copy);
        if (e == cudaSuccess) cudaEventElapsedTime(&me.h2d_ms, dv.ev[0], dv.ev[1]);
        if (e != cudaSuccess) return bail(fail(FMGI_ERR_CUDA, std::string("atlas upload: ") + cudaGetErrorString(e)));
    });

    int rc = FMGI_OK;
    for (int g = 0; g < G; g++)
        if (gpus[g].rc) { rc = gpus[g].rc; g_last_error = gpus[g].err; }

    // ---- phase 2: reduce-scatter fold over peer memory, then every GPU returns its slice -------------------------
    const double tf0 = now_ms();
    if (rc == FMGI_OK) {
        parallel_for(G, [&](int g) {
            PerGpu &me = gpus[g];
            auto bail = [&](cudaError_t e, const char *what) {
                me.rc = fail(FMGI_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); me.err = g_last_error;
            };
            cudaSetDevice(me.device);
            GpuRuntime::Dev &dv = rt.dev[me.device & 63];
            MemPool &pool = MemPool::get();
            const size_t slice = me.hi - me.lo;
            cudaError_t e = cudaEventRecord(dv.ev[2], dv.trace);
            if (slice) {
                std::vector<const float4 *> peers;
                size_t staged_at = 0;
                for (int q = 0; q < G && e == cudaSuccess; q++) {
                    if (q == g) continue;
                    if (rt.map_peer(me.device, gpus[q].device)) { peers.push_back(gpus[q].atlas + me.lo); continue; }
                    // no peer mapping: copy the peer's slice over
                    if (!me.staged) e = pool.alloc((void **)&me.staged, (size_t)(G - 1) * slice * sizeof(float4), false);
                    if (e == cudaSuccess)
                        e = cudaMemcpyPeerAsync(me.staged + staged_at, me.device, gpus[q].atlas + me.lo, gpus[q].device,
                                                slice * sizeof(float4), dv.trace);
                    peers.push_back(me.staged + staged_at);
                    staged_at += slice;
                }
                if (e == cudaSuccess) {
                    const int blocks = me.scene->num_sms * 8;
                    PeerList pl;
                    for (size_t q = 0; q < peers.size(); q++) pl.p[q] = peers[q];
                    k_fold_slice<<<blocks, 256, 0, dv.trace>>>(me.atlas + me.lo, me.init, pl, (int)peers.size(), slice);
                    me.scene->launches++;
                    e = cudaGetLastError();
                }
            }
            if (e == cudaSuccess) e = cudaEventRecord(dv.ev[3], dv.trace);
            if (e == cudaSuccess) e = cudaEventRecord(dv.ev[4], dv.trace);
            if (e == cudaSuccess && !tiles_out && slice)
                e = cudaMemcpyAsync(geo->texels + 4 * me.lo, me.atlas + me.lo, slice * sizeof(float4),
                                    cudaMemcpyDeviceToHost, dv.trace);
            if (e == cudaSuccess) e = cudaEventRecord(dv.ev[5], dv.trace);
            if (e == cudaSuccess) e = cudaStreamSynchronize(dv.trace);
            if (e != cudaSuccess) return bail(e, "atlas fold / read-back");
            cudaEventElapsedTime(&me.fold_ms, dv.ev[2], dv.ev[3]);
            cudaEventElapsedTime(&me.d2h_ms, dv.ev[4], dv.ev[5]);
        });
        for (int g = 0; g < G; g++)
            if (gpus[g].rc) { rc = gpus[g].rc; g_last_error = gpus[g].err; }
    }
    const double fold_host_ms = now_ms() - tf0;

    double tiles_ms = 0;
    if (rc == FMGI_OK && tiles_out) {
        // gather the folded slices on GPU 0, then normalise + tone-map + pack there: 3 bytes per texel come back
        const double t0 = now_ms();
        PerGpu &g0 = gpus[0];
        cudaSetDevice(g0.device);
        GpuRuntime::Dev &dv = rt.dev[g0.device & 63];
        cudaError_t e = cudaSuccess;
        for (int q = 1; q < G && e == cudaSuccess; q++)
            if (gpus[q].hi > gpus[q].lo)
                e = cudaMemcpyPeerAsync(g0.atlas + gpus[q].lo, g0.device, gpus[q].atlas + gpus[q].lo, gpus[q].device,
                                        (gpus[q].hi - gpus[q].lo) * sizeof(float4), dv.trace);
        const uint64_t tile_bytes = tiles_png ? fmgi_tile_png_bytes(geo->walls, geo->numWalls, nullptr)
                                              : fmgi_tile_bytes(geo->walls, geo->numWalls);
        unsigned char *d_rgb = nullptr;
        if (e == cudaSuccess) e = MemPool::get().alloc((void **)&d_rgb, tile_bytes ? tile_bytes : 16, false);
        if (e == cudaSuccess) {
            if (tonemap_impl(g0.scene, g0.atlas, spa, tint_extra, d_rgb, dv.trace, tiles_png) != FMGI_OK) e = cudaErrorUnknown;
            cudaSetDevice(g0.device);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(tiles_out, d_rgb, tile_bytes, cudaMemcpyDeviceToHost, dv.trace);
        if (e == cudaSuccess) e = cudaStreamSynchronize(dv.trace);
        MemPool::get().free(d_rgb);
        if (e != cudaSuccess) rc = fail(FMGI_ERR_CUDA, std::string("tile tone-map / read-back: ") + cudaGetErrorString(e));
        tiles_ms = now_ms() - t0;
    }

    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int g = 0; g < G; g++) {
            const fmgi_stats &s = gpus[g].st;
            stats->photons += s.photons; stats->rays += s.rays; stats->deposits += s.deposits;
            stats->mirror_bounces += s.mirror_bounces; stats->rect_tests += s.rect_tests;
            if (gpus[g].scene) stats->kernel_launches += gpus[g].scene->launches;
            stats->trace_ms = std::max(stats->trace_ms, s.trace_ms);
            stats->h2d_ms = std::max(stats->h2d_ms, (double)gpus[g].h2d_ms);
            stats->reduce_ms = std::max(stats->reduce_ms, (double)gpus[g].fold_ms);
            stats->d2h_ms = std::max(stats->d2h_ms, (double)gpus[g].d2h_ms);
            stats->init_ms = std::max(stats->init_ms, driver_ms + context_ms + gpus[g].init_ms);
            stats->upload_ms = std::max(stats->upload_ms, gpus[g].create_ms);
        }
        if (tiles_out) stats->d2h_ms = tiles_ms;
        stats->prepare_ms = build->prepare_ms;
        stats->grid_build_ms = build->grid_ms;
        for (int g = 0; g < G; g++)
            if (gpus[g].scene) stats->grid_build_ms = std::max(stats->grid_build_ms, build->grid_ms + gpus[g].scene->grid_device_ms);
        stats->num_gpus = G;
        stats->pool_rays = gpus[0].st.pool_rays;
        stats->bounds_violations = gpus[0].st.bounds_violations < 0 ? -1 : 0;
        for (int g = 0; g < G && stats->bounds_violations >= 0; g++) stats->bounds_violations += gpus[g].st.bounds_violations;
        stats->tier = gpus[0].st.tier;
        stats->num_sms = gpus[0].st.num_sms;
        stats->sm_clock_khz = gpus[0].st.sm_clock_khz;
    }
    for (int g = 0; g < G; g++) {
        PerGpu &me = gpus[g];
        cudaSetDevice(me.device);
        if (rt.dev[me.device & 63].ready) {
            cudaStreamSynchronize(rt.dev[me.device & 63].trace);
            cudaStreamSynchronize(rt.dev[me.device & 63].copy);
        }
        fmgi_scene_destroy(me.scene);
        MemPool &pool = MemPool::get();
        pool.free(me.atlas); pool.free(me.init); pool.free(me.staged);
    }
    // The reference entry point is one-shot (global_illumination_cl.c:315-320 releases everything): atlas-sized
    // blocks do not outlive it, only small tables and staging buffers stay cached (256 MB at most).  fmgi_bake /
    // fmgi_bake_tiles are called repeatedly by harnesses: they keep up to FMGI_CACHE_MB (default 8 GB, about three
    // atlases of the largest size the reference allows) so that a bake does not pay cudaMalloc / cudaFree of
    // gigabyte blocks every call - measured up to 0.7 s per call when the driver has to unmap them;
    // fmgi_release_cache() returns everything.
    size_t keep_mb = one_shot ? 256 : 8192;
    if (const char *v = getenv("FMGI_CACHE_MB")) keep_mb = (size_t)strtoull(v, nullptr, 0);
    const double tt0 = now_ms();
    MemPool::get().trim(keep_mb << 20);
    const double trim_ms = now_ms() - tt0;
    if (stats) stats->total_ms = now_ms() - t_begin;
    if (getenv("FMGI_DEBUG_TIMING"))
        fprintf(stderr, "[fmgi] bake: total %.3f ms (host tables %.3f [grid %.3f], context/streams %.3f, table upload %.3f, "
                        "atlas h2d %.3f (overlapped), trace+sync host %.3f [device %.3f], fold %.3f + d2h %.3f [host %.3f], "
                        "tiles %.3f, atlas alloc %.3f, cache trim %.3f, driver init %.3f, context %.3f; trace enqueue %.3f, h2d enqueue %.3f)\n",
                now_ms() - t_begin, build_ms, build->grid_ms, gpus[0].init_ms, gpus[0].create_ms, gpus[0].h2d_ms,
                gpus[0].sync_ms, gpus[0].st.trace_ms, gpus[0].fold_ms, gpus[0].d2h_ms, fold_host_ms, tiles_ms,
                gpus[0].alloc_ms, trim_ms, driver_ms, context_ms, gpus[0].trace_call_ms, gpus[0].h2d_call_ms);
    return rc;
}

}  // namespace

extern "C" {

int fmgi_bake(struct Geometry *geo, int spa, const fmgi_options *opt, fmgi_stats *stats)
{
    return bake_impl(geo, spa, opt, stats, nullptr, 0);
}

int fmgi_bake_tiles(struct Geometry *geo, int spa, const fmgi_options *opt, int tint_extra, uint8_t *rgb_out,
                    fmgi_stats *stats)
{
    if (!rgb_out) return fail(FMGI_ERR_ARG, "rgb_out is NULL");
    return bake_impl(geo, spa, opt, stats, rgb_out, tint_extra);
}

int fmgi_bake_tiles_png(struct Geometry *geo, int spa, const fmgi_options *opt, int tint_extra, uint8_t *png_out,
                        fmgi_stats *stats)
{
    if (!png_out) return fail(FMGI_ERR_ARG, "png_out is NULL");
    return bake_impl(geo, spa, opt, stats, png_out, tint_extra, false, true);
}

// ---- the reference boundary (global_illumination_cl.h:10) -------------------------------------------------

void performGlobalIlluminationCl(struct Geometry *geo, int numSamplesPerArea)
{
    fmgi_options o;
    fmgi_default_options(&o);
    if (const char *v = getenv("FMGI_MAX_DEPTH")) o.max_depth = atoi(v);
    if (const char *v = getenv("FMGI_SEED")) o.seed = (uint32_t)strtoul(v, nullptr, 0);
    if (const char *v = getenv("FMGI_GPUS")) o.num_gpus = atoi(v);
    if (const char *v = getenv("FMGI_DEPOSIT")) o.deposit = atoi(v);
    fmgi_stats st;
    const int rc = bake_impl(geo, numSamplesPerArea, &o, &st, nullptr, 0, true);
    if (rc != FMGI_OK) {
        printf("[Err] photon mapping on the GPU failed: %s\n", fmgi_last_error());
        exit(1);
    }
    printf("[INF] photon-mapped %llu photons / %llu bounces on %d GPU(s): trace %.1f ms (%.3g bounces/s), total %.1f ms\n",
           (unsigned long long)st.photons, (unsigned long long)st.deposits, st.num_gpus, st.trace_ms,
           st.trace_ms > 0 ? st.deposits / (st.trace_ms * 1e-3) : 0.0, st.total_ms);
    int verbose = 0;
    if (const char *v = getenv("FMGI_STATS")) verbose = atoi(v);
    if (verbose)
        printf("[INF] rays %llu, mirror bounces %llu, rectangle tests %llu, h2d %.2f ms, d2h %.2f ms, fold %.2f ms, "
               "launches %llu\n",
               (unsigned long long)st.rays, (unsigned long long)st.mirror_bounces,
               (unsigned long long)st.rect_tests, st.h2d_ms, st.d2h_ms, st.reduce_ms,
               (unsigned long long)st.kernel_launches);
    if (verbose >= 2) {
        // when the caller's own exit begins (tiles written, geometry freed): what follows is process teardown
        atexit([] { printf("[INF] fmgi exit_begin %.1f ms\n", ms_since_process_start()); fflush(stdout); });
    }
    if (verbose >= 2)
        // where the call's time went; before_call = process start -> this call (loader, parseLayout), from /proc
        printf("[INF] fmgi breakdown: before_call %.1f ms, init %.1f ms, tables %.2f ms, upload %.2f ms, trace %.2f ms, "
               "fold %.2f ms, d2h %.2f ms, call %.1f ms\n",
               ms_since_process_start() - st.total_ms, st.init_ms, st.prepare_ms + st.grid_build_ms, st.upload_ms,
               st.trace_ms, st.reduce_ms, st.d2h_ms, st.total_ms);
}

// ---- tile post-processing ----------------------------------------------------------------

uint64_t fmgi_tile_bytes(const fmgi_rect *walls, int num_walls)
{
    uint64_t n = 0;
    for (int i = 0; walls && i < num_walls; i++) n += 3ull * (uint64_t)walls[i].lightmap[1] * (uint64_t)walls[i].lightmap[2];
    return n;
}

} // extern "C"

namespace {

// Device tone-map of every wall's base level: packed RGB (png == false) or complete PNG files (png == true).
int tonemap_impl(fmgi_scene *s, const void *atlas_dev, int spa, int tint_extra, void *out_dev, cudaStream_t st, bool png)
{
    if (!s || !atlas_dev || !out_dev) return fail(FMGI_ERR_ARG, "scene, atlas or output is NULL");
    DeviceGuard guard(s->device);
    const int W = s->host.num_walls;
    MemPool &pool = MemPool::get();
    if (!s->d_tile_walls) {
        FMGI_CUDA(pool.alloc((void **)&s->d_tile_walls, (W ? W : 1) * sizeof(TileWall), false));
        FMGI_CUDA(pool.alloc((void **)&s->h_tile_walls, (W ? W : 1) * sizeof(TileWall), true));
        FMGI_CUDA(pool.alloc((void **)&s->d_png_walls, (W ? W : 1) * sizeof(PngWall), false));
        FMGI_CUDA(pool.alloc((void **)&s->h_png_walls, (W ? W : 1) * sizeof(PngWall), true));
    } else {
        FMGI_CUDA(cudaStreamSynchronize(st));                     // pinned staging is reused
    }
    long long pixels = 0, file_off = 0;
    for (int i = 0; i < W; i++) {
        const ShadeRect &sh = s->host.shade[i];
        const int tw = sh.tiles & 0xffff, th = sh.tiles >> 16;
        const int tiles = tw * th;
        TileWall &t = s->h_tile_walls[i];
        t.base = sh.base;
        t.first = (int32_t)pixels;
        const float tiles_per_sample = tiles / (s->build->wall_area[i] * spa);            // main.c:73 (int / (float * int))
        t.scale = (float)(0.35 * tiles_per_sample);                                // main.c:77: double product, float arg
        t.is_floor = s->build->wall_floor[i];
        pixels += tiles;
        PngWall &pw = s->h_png_walls[i];
        pw.file_off = file_off; pw.width = tw; pw.height = th;
        file_off += (long long)png_file_bytes(tw, th);
    }
    if (pixels > 0x7fffffffLL) return fail(FMGI_ERR_UNSUPPORTED, "more than 2^31 tile pixels");
    if (pixels == 0) return FMGI_OK;
    FMGI_CUDA(cudaMemcpyAsync(s->d_tile_walls, s->h_tile_walls, W * sizeof(TileWall), cudaMemcpyHostToDevice, st));
    long long blocks = (pixels + 255) / 256;
    if (blocks > (long long)s->num_sms * 16) blocks = (long long)s->num_sms * 16;
    if (!png) {
        k_tonemap<false><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(atlas_dev), s->d_tile_walls, W, pixels,
                                                    tint_extra, reinterpret_cast<unsigned char *>(out_dev), nullptr);
        s->launches++;
    } else {
        FMGI_CUDA(cudaMemcpyAsync(s->d_png_walls, s->h_png_walls, W * sizeof(PngWall), cudaMemcpyHostToDevice, st));
        k_tonemap<true><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(atlas_dev), s->d_tile_walls, W, pixels,
                                                   tint_extra, reinterpret_cast<unsigned char *>(out_dev), s->d_png_walls);
        k_png_finish<<<(W + 255) / 256, 256, 0, st>>>(s->d_png_walls, W, reinterpret_cast<unsigned char *>(out_dev));
        s->launches += 2;
    }
    s->note_stream(st);
    FMGI_CUDA(cudaGetLastError());
    return FMGI_OK;
}

}  // namespace

extern "C" {

int fmgi_scene_tonemap(fmgi_scene *s, const void *atlas_dev, int spa, int tint_extra, void *rgb_dev, void *cuda_stream)
{
    return tonemap_impl(s, atlas_dev, spa, tint_extra, rgb_dev, (cudaStream_t)cuda_stream, false);
}

uint64_t fmgi_tile_png_bytes(const fmgi_rect *walls, int num_walls, uint64_t *offsets_out)
{
    uint64_t n = 0;
    for (int i = 0; walls && i < num_walls; i++) {
        if (offsets_out) offsets_out[i] = n;
        n += png_file_bytes(walls[i].lightmap[1], walls[i].lightmap[2]);
    }
    if (offsets_out && walls) offsets_out[num_walls] = n;
    return n;
}

int fmgi_scene_tiles_png(fmgi_scene *s, const void *atlas_dev, int spa, int tint_extra, void *png_dev, void *cuda_stream)
{
    return tonemap_impl(s, atlas_dev, spa, tint_extra, png_dev, (cudaStream_t)cuda_stream, true);
}

// ---- ambient occlusion --------------------------------------------------------------------

int fmgi_geosphere(int iterations, float *xyz_out, int max_directions)
{
    const std::vector<float> d = geosphere_directions(iterations);
    const int n = (int)d.size() / 3;
    for (int i = 0; xyz_out && i < 3 * (n < max_directions ? n : max_directions); i++) xyz_out[i] = d[i];
    return n;
}

int fmgi_scene_ambient_occlusion(fmgi_scene *s, void *atlas_dev, void *cuda_stream)
{
    if (!s || !atlas_dev) return fail(FMGI_ERR_ARG, "scene or atlas is NULL");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int W = s->host.num_walls;
    // flat enumeration of the base-level texels
    std::vector<AoWall> aw(W ? W : 1);
    long long pixels = 0;
    for (int i = 0; i < W; i++) {
        const ShadeRect &sh = s->host.shade[i];
        aw[i].base = sh.base; aw[i].first = (int32_t)pixels; aw[i].wall = i; aw[i].pad = 0;
        pixels += (long long)(sh.tiles & 0xffff) * (sh.tiles >> 16);
    }
    if (pixels > 0x7fffffffLL) return fail(FMGI_ERR_UNSUPPORTED, "more than 2^31 texels");
    if (pixels == 0) return FMGI_OK;
    // the reference's geoSphere4 direction set (photonmap.c:450-453), regenerated (geosphere.cpp)
    const std::vector<float> dirs = geosphere_directions(4);
    const int num_dirs = (int)dirs.size() / 3;
    std::vector<float> blob;                                   // widths | heights | float4 directions
    blob.reserve((size_t)6 * W + 4 * num_dirs + 4);
    for (int i = 0; i < W; i++) for (int c = 0; c < 3; c++) blob.push_back(s->build->wall_wh[6 * i + c]);
    for (int i = 0; i < W; i++) for (int c = 0; c < 3; c++) blob.push_back(s->build->wall_wh[6 * i + 3 + c]);
    while (blob.size() % 4) blob.push_back(0.0f);
    const size_t dirs_at = blob.size();
    for (int k = 0; k < num_dirs; k++) {
        blob.push_back(dirs[3 * k]); blob.push_back(dirs[3 * k + 1]); blob.push_back(dirs[3 * k + 2]); blob.push_back(0.0f);
    }
    MemPool &pool = MemPool::get();
    if (!s->d_ao) {
        FMGI_CUDA(pool.alloc((void **)&s->d_ao, blob.size() * sizeof(float), false));
        FMGI_CUDA(pool.alloc((void **)&s->d_ao_walls, aw.size() * sizeof(AoWall), false));
    }
    FMGI_CUDA(cudaMemcpyAsync(s->d_ao, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    FMGI_CUDA(cudaMemcpyAsync(s->d_ao_walls, aw.data(), aw.size() * sizeof(AoWall), cudaMemcpyHostToDevice, st));
    FMGI_CUDA(cudaStreamSynchronize(st));                      // the staging vectors die with this call

    TraceParams p = base_params(s);
    p.atlas = reinterpret_cast<float4 *>(atlas_dev);
    p.ao_width = s->d_ao;
    p.ao_height = s->d_ao + 3 * (size_t)W;
    const float4 *d_dirs = reinterpret_cast<const float4 *>(s->d_ao + dirs_at);
    long long blocks = (pixels + kTraceThreads - 1) / kTraceThreads;
    if (blocks > (long long)s->num_sms * 8) blocks = (long long)s->num_sms * 8;
    auto go = [&](auto kernel) {
        cudaError_t e = cudaSuccess;
        if (s->smem_bytes > 48 * 1024)
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes);
        if (e != cudaSuccess) return e;
        kernel<<<(int)blocks, kTraceThreads, s->smem_bytes, st>>>(p, s->d_ao_walls, W, pixels, d_dirs, num_dirs);
        s->launches++;
        return cudaGetLastError();
    };
    if (s->kernel_tier == kTierRooms) FMGI_CUDA(go(k_ambient_occlusion<kTierRooms>));
    else if (s->kernel_tier == kTierSoupPlanes) FMGI_CUDA(go(k_ambient_occlusion<kTierSoupPlanes>));
    else if (s->kernel_tier == FMGI_TIER_SOUP) FMGI_CUDA(go(k_ambient_occlusion<FMGI_TIER_SOUP>));
    else FMGI_CUDA(go(k_ambient_occlusion<FMGI_TIER_GRID>));
    s->note_stream(st);
    return FMGI_OK;
}

int fmgi_ambient_occlusion(struct Geometry *geo_, const fmgi_options *opt)
{
    fmgi_geometry *geo = reinterpret_cast<fmgi_geometry *>(geo_);
    if (!geo || geo->numTexels < 0 || (geo->numTexels && !geo->texels)) return fail(FMGI_ERR_ARG, "bad geometry");
    fmgi_options o = resolve(opt);
    fmgi_scene *scene = nullptr;
    int rc = fmgi_scene_create(&scene, geo->walls, geo->numWalls, geo->windows, geo->numWindows, geo->lights,
                               geo->numLights, geo->numTexels, &o);
    if (rc) return rc;
    DeviceGuard guard(o.device);
    const size_t bytes = (size_t)geo->numTexels * sizeof(float4);
    float4 *atlas = nullptr;
    cudaError_t e = MemPool::get().alloc((void **)&atlas, bytes ? bytes : 16, false);
    // texels the pass does not write (mip slots) keep their contents, as in the reference
    if (e == cudaSuccess) e = cudaMemcpy(atlas, geo->texels, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = fmgi_scene_ambient_occlusion(scene, atlas, nullptr);
        cudaSetDevice(o.device);
        if (rc == FMGI_OK) e = cudaMemcpy(geo->texels, atlas, bytes, cudaMemcpyDeviceToHost);
    }
    MemPool::get().free(atlas);
    fmgi_scene_destroy(scene);
    {
        size_t keep_mb = 8192;                     // as fmgi_bake: bounded cache per device
        if (const char *v = getenv("FMGI_CACHE_MB")) keep_mb = (size_t)strtoull(v, nullptr, 0);
        MemPool::get().trim(keep_mb << 20);
    }
    if (rc == FMGI_OK && e != cudaSuccess) rc = fail(FMGI_ERR_CUDA, std::string("ambient occlusion: ") + cudaGetErrorString(e));
    return rc;
}

// ---- parity probes ---------------------------------------------------------------------------------------------

int fmgi_probe_closest_hit(fmgi_scene *s, const float *origins, const float *dirs, int n, int32_t *hit_index,
                           float *hit_dist)
{
    if (!s || !origins || !dirs || !hit_index || !hit_dist || n < 0) return fail(FMGI_ERR_ARG, "bad argument");
    if (n == 0) return FMGI_OK;
    DeviceGuard guard(s->device);
    DevBuf<float> d_o, d_d, d_t;
    DevBuf<int32_t> d_i;
    const size_t vb = (size_t)n * 3 * sizeof(float);
    FMGI_CUDA(d_o.alloc((size_t)n * 3));
    FMGI_CUDA(d_d.alloc((size_t)n * 3));
    FMGI_CUDA(d_t.alloc((size_t)n));
    FMGI_CUDA(d_i.alloc((size_t)n));
    FMGI_CUDA(cudaMemcpy(d_o, origins, vb, cudaMemcpyHostToDevice));
    FMGI_CUDA(cudaMemcpy(d_d, dirs, vb, cudaMemcpyHostToDevice));
    const TraceParams p = base_params(s);
    int blocks = (n + 255) / 256;
    if (blocks > s->num_sms * 4) blocks = s->num_sms * 4;
    if (s->kernel_tier == kTierRooms) {
        k_probe_closest_hit<kTierRooms><<<blocks, 256>>>(p, d_o, d_d, n, d_i, d_t);
    } else if (s->kernel_tier == kTierSoupPlanes) {
        if (s->smem_bytes > 48 * 1024)
            FMGI_CUDA(cudaFuncSetAttribute(k_probe_closest_hit<kTierSoupPlanes>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        k_probe_closest_hit<kTierSoupPlanes><<<blocks, 256, s->smem_bytes>>>(p, d_o, d_d, n, d_i, d_t);
    } else if (s->tier == FMGI_TIER_SOUP) {
        if (s->smem_bytes > 48 * 1024)
            FMGI_CUDA(cudaFuncSetAttribute(k_probe_closest_hit<FMGI_TIER_SOUP>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        k_probe_closest_hit<FMGI_TIER_SOUP><<<blocks, 256, s->smem_bytes>>>(p, d_o, d_d, n, d_i, d_t);
    } else {
        k_probe_closest_hit<FMGI_TIER_GRID><<<blocks, 256>>>(p, d_o, d_d, n, d_i, d_t);
    }
    s->launches++;
    FMGI_CUDA(cudaGetLastError());
    FMGI_CUDA(cudaMemcpy(hit_index, d_i, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    FMGI_CUDA(cudaMemcpy(hit_dist, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return FMGI_OK;
}

int fmgi_probe_tile_ids(fmgi_scene *s, const int32_t *rect_index, const float *points, int n, int32_t *tile_ids)
{
    if (!s || !rect_index || !points || !tile_ids || n < 0) return fail(FMGI_ERR_ARG, "bad argument");
    if (n == 0) return FMGI_OK;
    for (int i = 0; i < n; i++)
        if (rect_index[i] < 0 || rect_index[i] >= s->host.num_walls) return fail(FMGI_ERR_ARG, "rect index out of range");
    DeviceGuard guard(s->device);
    DevBuf<float> d_p;
    DevBuf<int32_t> d_r, d_t;
    FMGI_CUDA(d_p.alloc((size_t)n * 3));
    FMGI_CUDA(d_r.alloc((size_t)n));
    FMGI_CUDA(d_t.alloc((size_t)n));
    FMGI_CUDA(cudaMemcpy(d_p, points, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    FMGI_CUDA(cudaMemcpy(d_r, rect_index, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    int blocks = (n + 255) / 256;
    if (blocks > s->num_sms * 8) blocks = s->num_sms * 8;
    k_probe_tile_ids<<<blocks, 256>>>(reinterpret_cast<const float4 *>(s->d_shade), d_r, d_p, n, d_t);
    s->launches++;
    FMGI_CUDA(cudaGetLastError());
    FMGI_CUDA(cudaMemcpy(tile_ids, d_t, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return FMGI_OK;
}

int fmgi_probe_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    if (!ctr || !key || !out) return fail(FMGI_ERR_ARG, "bad argument");
    DevBuf<uint32_t> d;
    FMGI_CUDA(d.alloc(4));
    k_probe_philox<<<1, 1>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], d);
    FMGI_CUDA(cudaGetLastError());
    FMGI_CUDA(cudaMemcpy(out, d, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return FMGI_OK;
}

int64_t fmgi_probe_grid_table(fmgi_scene *s, void *records_out, uint64_t max_records)
{
    if (!s) { fail(FMGI_ERR_ARG, "scene is NULL"); return -1; }
    if (records_out && max_records) {
        DeviceGuard guard(s->device);
        const size_t n = s->grid_records < max_records ? s->grid_records : (size_t)max_records;
        if (n && cudaMemcpy(records_out, s->d_grid_table, n * sizeof(GridRec), cudaMemcpyDeviceToHost) != cudaSuccess) {
            fail(FMGI_ERR_CUDA, "grid table read-back failed");
            return -1;
        }
    }
    return (int64_t)s->grid_records;
}

int fmgi_probe_philox2x32(const uint32_t ctr[2], uint32_t key, uint32_t out[2])
{
    if (!ctr || !out) return fail(FMGI_ERR_ARG, "bad argument");
    DevBuf<uint32_t> d;
    FMGI_CUDA(d.alloc(2));
    k_probe_philox2<<<1, 1>>>(ctr[0], ctr[1], key, d);
    FMGI_CUDA(cudaGetLastError());
    FMGI_CUDA(cudaMemcpy(out, d, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return FMGI_OK;
}

int fmgi_probe_deposit_peak(uint64_t num_texels, uint64_t num_deposits, int device, double *deposits_per_s)
{
    if (!deposits_per_s || num_texels == 0 || num_texels > 0xffffffffull || num_deposits == 0)
        return fail(FMGI_ERR_ARG, "bad argument");
    int ndev = 0;
    FMGI_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(FMGI_ERR_ARG, "device ordinal out of range");
    DeviceGuard guard(device);
    int sms = 0;
    FMGI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    DevBuf<float4> atlas;
    FMGI_CUDA(atlas.alloc((size_t)num_texels));
    FMGI_CUDA(cudaMemset(atlas, 0, (size_t)num_texels * sizeof(float4)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    FMGI_CUDA(cudaEventCreate(&e0));
    cudaError_t err = cudaEventCreate(&e1);
    float ms = 0;
    if (err == cudaSuccess) {
        k_probe_red_peak<<<sms * 8, 256>>>(atlas, (uint32_t)num_texels, num_deposits / 4 + 1, 1u);      // warm-up
        cudaEventRecord(e0);
        k_probe_red_peak<<<sms * 8, 256>>>(atlas, (uint32_t)num_texels, num_deposits, 2u);
        cudaEventRecord(e1);
        err = cudaEventSynchronize(e1);
        if (err == cudaSuccess) err = cudaGetLastError();
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (err != cudaSuccess) return fail(FMGI_ERR_CUDA, std::string("deposit peak probe: ") + cudaGetErrorString(err));
    *deposits_per_s = ms > 0 ? (double)num_deposits / (ms * 1e-3) : 0.0;
    return FMGI_OK;
}

int fmgi_probe_sample_dirs(const float normal[3], int sky, uint32_t seed, int n, float *dirs_out)
{
    if (!normal || !dirs_out || n < 0) return fail(FMGI_ERR_ARG, "bad argument");
    if (n == 0) return FMGI_OK;
    float u[3], v[3];
    sampler_basis(normal, u, v);
    DevBuf<float> d;
    FMGI_CUDA(d.alloc((size_t)n * 3));
    k_probe_sample_dirs<<<(n + 255) / 256, 256>>>(make_float4(normal[0], normal[1], normal[2], 0),
                                                  make_float4(u[0], u[1], u[2], 0), make_float4(v[0], v[1], v[2], 0),
                                                  sky, seed, n, d);
    FMGI_CUDA(cudaGetLastError());
    FMGI_CUDA(cudaMemcpy(dirs_out, d, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return FMGI_OK;
}

int fmgi_probe_paths(fmgi_scene *s, int emitter_index, int max_depth, uint32_t seed, uint64_t first, int count,
                     int32_t *texel_out)
{
    if (!s || !texel_out || count < 0 || max_depth < 1 || max_depth > kPhiloxMaxDepth) return fail(FMGI_ERR_ARG, "bad argument");
    const int E = (int)s->host.emitters.size();
    if (emitter_index < 0 || emitter_index >= E) return fail(FMGI_ERR_ARG, "emitter index out of range");
    if (count == 0) return FMGI_OK;
    DeviceGuard guard(s->device);
    if (s->traced) FMGI_CUDA(cudaEventSynchronize(s->ev_stop));
    // one-emitter job table: every other emitter has an empty range
    unsigned long long total = 0;
    for (int e = 0; e < E; e++) {
        s->h_jobs[e] = total;
        s->h_jobs[E + 1 + e] = e == emitter_index ? first : 0;
        s->h_jobs[2 * E + 1 + e] = e == emitter_index ? (unsigned long long)count : 0;
        if (e == emitter_index) total += ((unsigned long long)count + kMinChunkPhotons - 1) / kMinChunkPhotons;
    }
    s->h_jobs[E] = total;
    FMGI_CUDA(cudaMemcpy(s->d_jobs, s->h_jobs, job_table_words(E) * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    FMGI_CUDA(cudaMemset(s->d_counters, 0, 8 * sizeof(unsigned long long)));
    DevBuf<int32_t> d_path;
    const size_t pb = (size_t)count * max_depth * sizeof(int32_t);
    FMGI_CUDA(d_path.alloc((size_t)count * max_depth));
    FMGI_CUDA(cudaMemset(d_path, 0xff, pb));
    TraceParams p = base_params(s);
    p.total_jobs = total;
    p.chunk = kMinChunkPhotons;
    p.max_depth = max_depth;
    set_seed(p, seed);
    p.path_out = d_path;
    int blocks = (count + kTraceThreads - 1) / kTraceThreads;
    if (blocks > s->num_sms) blocks = s->num_sms;
    FMGI_CUDA(launch_trace(s, p, 0, true, blocks, nullptr));
    FMGI_CUDA(cudaMemcpy(texel_out, d_path, pb, cudaMemcpyDeviceToHost));
    return FMGI_OK;
}

}  // extern "C"
