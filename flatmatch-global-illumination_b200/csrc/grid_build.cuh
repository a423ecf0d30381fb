// Device-side assembly of the floor-plan grid table T (scene_tables.h) - the per-cell half of build_grid
// (scene_prep.cpp: grid_assemble_host), which replaces the reference's O(n^2)-per-node BSP build
// (photonmap.c:302-374) for this path.  The host classifies the colliders (grid_classify: O(n), 1-2 ms for
// 21.5k rectangles); binning them into 10^5 (list, cell) lists, ordering each list and laying out T is
// data-parallel and runs here:
//   k_grid_count   one thread per collider: count[list, cell]++ over the cells it covers
//   exclusive scans (cub) of count -> begin, and of max(count - 1, 0) -> where a list's rest records go
//   k_grid_fill    one thread per collider: scatter {record, wall order key} into the lists (atomic slot claim)
//   k_grid_layout  one thread per (list, cell): order the list the way the host does - walk lists by wall
//                  index, plane lists by the area of the cell the rectangle covers, largest first, ties by
//                  wall index - then write the head (first record inline + continuation range) and the rest
// so that the table is bit-identical to the host-built one (tested: test_device_grid_table_equals_host).
#pragma once
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include "mem_pool.h"
#include "scene_tables.h"

namespace fmgi {

struct GridBuildParams {
    GridDesc g;
    int num_items;
    int num_used;                 // lists in T: planes_up + planes_down + 4
};

// compact list number (position in T) of plane list `list` (CSR numbering: < 8 up planes, 8.. down planes)
__host__ __device__ __forceinline__ int grid_used_list(const GridDesc &g, int list)
{
    return list < kMaxPlanesPerSign ? list : g.planes_up + (list - kMaxPlanesPerSign);
}

// Calls fn(compact list, cell) for every (list, cell) the item goes to - the order of grid_assemble_host.
template <typename Fn>
__device__ __forceinline__ void grid_for_each_slot(const GridBuildParams &bp, const GridItem &it, Fn fn)
{
    const GridDesc &g = bp.g;
    const int walk0 = g.planes_up + g.planes_down;
    for (int c = 0; c < (it.list >= 0 ? 1 : 4); c++) {
        int u;
        if (it.list >= 0) u = grid_used_list(g, it.list);
        else {
            // normal +x is faced by d.x < 0 (combo bit 0 clear), normal -x by d.x > 0; same for y
            if (it.axis == 0 && ((c & 1) != 0) != (it.neg != 0)) continue;
            if (it.axis == 1 && ((c & 2) != 0) != (it.neg != 0)) continue;
            u = walk0 + c;
        }
        for (int cy = it.cy0; cy <= it.cy1; cy++)
            for (int cx = it.cx0; cx <= it.cx1; cx++) fn(u, cy * g.nx + cx);
    }
}

__global__ void k_grid_count(const GridBuildParams bp, const GridItem *__restrict__ items, int32_t *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bp.num_items) return;
    const GridItem it = items[i];
    grid_for_each_slot(bp, it, [&](int u, int cell) { atomicAdd(count + (size_t)u * bp.g.ncell + cell, 1); });
}

// rest[i] = records of list i that do not fit the head
__global__ void k_grid_rest(const int32_t *__restrict__ count, int32_t *__restrict__ rest, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rest[i] = count[i] > 0 ? count[i] - 1 : 0;
}

struct GridSlotRec {
    GridRec rec;
    int32_t order;                // index of the item (wall order)
    int32_t pad[7];
};
static_assert(sizeof(GridSlotRec) == 64, "GridSlotRec is 64 bytes");

__global__ void k_grid_fill(const GridBuildParams bp, const GridItem *__restrict__ items, const int32_t *__restrict__ begin,
                            int32_t *__restrict__ fill, GridSlotRec *__restrict__ recs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bp.num_items) return;
    const GridItem it = items[i];
    grid_for_each_slot(bp, it, [&](int u, int cell) {
        const size_t l = (size_t)u * bp.g.ncell + cell;
        const int pos = atomicAdd(fill + l, 1);
        GridSlotRec r;
        r.rec = it.rec;
        r.order = i;
        recs[begin[l] + pos] = r;
    });
}

// area of cell (cx, cy) the record's rectangle covers - grid_assemble_host's `overlap`, same float operations
__device__ __forceinline__ float grid_overlap(const GridDesc &g, const GridRec &r, int cx, int cy)
{
    const float x0c = __fadd_rn(g.x0, __fmul_rn((float)cx, g.cell)), y0c = __fadd_rn(g.y0, __fmul_rn((float)cy, g.cell));
    const float ox = __fsub_rn(fminf(__fadd_rn(r.mid_i, r.half_i), __fadd_rn(x0c, g.cell)), fmaxf(__fsub_rn(r.mid_i, r.half_i), x0c));
    const float oy = __fsub_rn(fminf(__fadd_rn(r.mid_j, r.half_j), __fadd_rn(y0c, g.cell)), fmaxf(__fsub_rn(r.mid_j, r.half_j), y0c));
    return __fmul_rn(fmaxf(ox, 0.0f), fmaxf(oy, 0.0f));
}

__global__ void k_grid_layout(const GridBuildParams bp, const int32_t *__restrict__ count, const int32_t *__restrict__ begin,
                              const int32_t *__restrict__ rest_begin, GridSlotRec *__restrict__ recs, GridRec *__restrict__ table)
{
    const GridDesc &g = bp.g;
    const size_t num_lists = (size_t)bp.num_used * g.ncell;
    const size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (l >= num_lists) return;
    const int n = count[l];
    GridRec head;
    head.mid_i = 0.0f; head.half_i = -1.0f; head.mid_j = 0.0f; head.half_j = -1.0f;      // the dummy no ray can hit
    head.c = __int_as_float(0x7fc00000); head.tag = 0u; head.next = 0; head.end = 0;
    if (n > 0) {
        GridSlotRec *r = recs + begin[l];
        const int u = (int)(l / g.ncell), cell = (int)(l % g.ncell);
        const bool plane = u < g.planes_up + g.planes_down;
        const int cx = cell % g.nx, cy = cell / g.nx;
        // insertion sort (lists hold a handful of records): plane lists by covered area, largest first, then wall
        // order; walk lists by wall order
        for (int a = 1; a < n; a++) {
            const GridSlotRec x = r[a];
            const float kx = plane ? grid_overlap(g, x.rec, cx, cy) : 0.0f;
            int b = a - 1;
            while (b >= 0) {
                const float kb = plane ? grid_overlap(g, r[b].rec, cx, cy) : 0.0f;
                const bool after = kb < kx || (kb == kx && r[b].order > x.order);     // r[b] belongs behind x
                if (!after) break;
                r[b + 1] = r[b];
                b--;
            }
            r[b + 1] = x;
        }
        const int first_rest = (int)num_lists + rest_begin[l];
        head = r[0].rec;
        head.next = first_rest;
        head.end = first_rest + (n - 1);
        for (int q = 1; q < n; q++) {
            GridRec rr = r[q].rec;
            rr.next = first_rest + q;              // what follows THIS record: the list's remaining records
            rr.end = head.end;
            table[first_rest + q - 1] = rr;
        }
    }
    table[l] = head;
}

// Builds T on the current device from the classified items.  *table_out: pool block of *records_out GridRec
// (heads first, then the rest records).  Asynchronous on `st` except for one 4-byte read-back of the rest total.
inline cudaError_t grid_assemble_device(const GridDesc &g, const std::vector<GridItem> &items, GridRec **table_out,
                                        size_t *records_out, cudaStream_t st)
{
    *table_out = nullptr;
    *records_out = 0;
    GridBuildParams bp;
    bp.g = g;
    bp.num_items = (int)items.size();
    bp.num_used = g.planes_up + g.planes_down + 4;
    const size_t num_lists = (size_t)bp.num_used * g.ncell;
    if (num_lists >= (1u << 30)) return cudaErrorInvalidValue;
    // entries of all lists (host side, exact): every item covers (cx1-cx0+1)*(cy1-cy0+1) cells of 1 / 2 / 4 lists
    size_t total = 0;
    for (const GridItem &it : items) {
        const size_t cells = (size_t)(it.cx1 - it.cx0 + 1) * (size_t)(it.cy1 - it.cy0 + 1);
        total += cells * (it.list >= 0 ? 1 : (it.axis <= 1 ? 2 : 4));
    }
    MemPool &pool = MemPool::get();
    GridItem *d_items = nullptr;
    int32_t *d_count = nullptr, *d_begin = nullptr, *d_rest = nullptr, *d_rest_begin = nullptr, *d_fill = nullptr;
    GridSlotRec *d_recs = nullptr;
    void *d_tmp = nullptr;
    GridRec *d_table = nullptr;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = pool.alloc(p, bytes ? bytes : 16, false); };
    alloc((void **)&d_items, items.size() * sizeof(GridItem));
    alloc((void **)&d_count, (num_lists + 1) * sizeof(int32_t));
    alloc((void **)&d_begin, (num_lists + 1) * sizeof(int32_t));
    alloc((void **)&d_rest, (num_lists + 1) * sizeof(int32_t));
    alloc((void **)&d_rest_begin, (num_lists + 1) * sizeof(int32_t));
    alloc((void **)&d_fill, (num_lists + 1) * sizeof(int32_t));
    alloc((void **)&d_recs, total * sizeof(GridSlotRec));
    // heads + at most one rest record per list entry
    alloc((void **)&d_table, (num_lists + total) * sizeof(GridRec));
    size_t tmp_bytes = 0;
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_count, d_begin, (int)num_lists + 1, st);
    alloc(&d_tmp, tmp_bytes);
    if (e == cudaSuccess && !items.empty())
        e = cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(GridItem), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, (num_lists + 1) * sizeof(int32_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_fill, 0, (num_lists + 1) * sizeof(int32_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_rest, 0, (num_lists + 1) * sizeof(int32_t), st);
    int32_t rest_total = 0;
    if (e == cudaSuccess) {
        const int tb = 256;
        if (!items.empty()) k_grid_count<<<((int)items.size() + tb - 1) / tb, tb, 0, st>>>(bp, d_items, d_count);
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_count, d_begin, (int)num_lists + 1, st);
        k_grid_rest<<<(int)((num_lists + tb - 1) / tb), tb, 0, st>>>(d_count, d_rest, (int)num_lists);
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_rest, d_rest_begin, (int)num_lists + 1, st);
        if (!items.empty()) k_grid_fill<<<((int)items.size() + tb - 1) / tb, tb, 0, st>>>(bp, d_items, d_begin, d_fill, d_recs);
        k_grid_layout<<<(int)((num_lists + tb - 1) / tb), tb, 0, st>>>(bp, d_count, d_begin, d_rest_begin, d_recs, d_table);
        e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(&rest_total, d_rest_begin + num_lists, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // the staging vector and the scratch blocks die here
    }
    pool.free(d_items); pool.free(d_count); pool.free(d_begin); pool.free(d_rest); pool.free(d_rest_begin);
    pool.free(d_fill); pool.free(d_recs); pool.free(d_tmp);
    if (e != cudaSuccess) { pool.free(d_table); return e; }
    *table_out = d_table;
    *records_out = num_lists + (size_t)rest_total;
    return cudaSuccess;
}

}  // namespace fmgi
