// grid_build.cuh -- everything between the raw point arrays and the sorted uniform grid:
//   world box + radius reduction      (replaces _update_world_AABB[_simd],  TreeNSearch.cpp:415-645)
//   cell assignment + Morton keys      (replaces _points_to_cells[_simd],    TreeNSearch.cpp:646-1113, and libmorton)
//   reorder gather into sorted float4  (replaces the leaf gather,            TreeNSearch.cpp:2161-2399)
//   cell start/end compaction + hash   (replaces CellList,                   internals/octree_internals.h:63-159)
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace tnsb {

// ----------------------------------------------------------------------------------------------------------------------
// World box / radius range.  out[0..2] = min xyz, out[3..5] = max xyz, out[6] = min radius, out[7] = max radius, all in
// the order preserving uint encoding (initialise mins with 0xffffffff and maxs with 0).
// For double input the (float) cast of the reference (TreeNSearch.cpp:275-296) happens here and the converted arrays are
// written out so that every later kernel only sees float.
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<double>(double v) { return __double2float_rn(v); }

constexpr int kAabbThreads = 256;

// `stride` = elements between consecutive points of the input (3 for xyzxyz, 4 for (x,y,z,id) records); the float
// conversion output is always packed xyz.
template <typename T>
__global__ void __launch_bounds__(kAabbThreads) aabb_kernel(const T* __restrict__ pts, const T* __restrict__ radii, int n, int stride,
                                                            float* __restrict__ pts_f32_out, float* __restrict__ radii_f32_out,
                                                            uint32_t* __restrict__ out)
{
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    float rlo = INFINITY, rhi = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = to_f32<T>(pts[i * stride + c]);
            if (pts_f32_out) pts_f32_out[i * 3 + c] = v;
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
    if (radii) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const float r = to_f32<T>(radii[i]);
            if (radii_f32_out) radii_f32_out[i] = r;
            rlo = fminf(rlo, r);
            rhi = fmaxf(rhi, r);
        }
    }
    uint32_t v[8] = { float_to_ordered(lo[0]), float_to_ordered(lo[1]), float_to_ordered(lo[2]),
                      float_to_ordered(hi[0]), float_to_ordered(hi[1]), float_to_ordered(hi[2]),
                      float_to_ordered(rlo), float_to_ordered(rhi) };
    __shared__ uint32_t s_red[8][kAabbThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const bool is_min = (k < 3) || (k == 6);
        const uint32_t r = is_min ? __reduce_min_sync(kFull, v[k]) : __reduce_max_sync(kFull, v[k]);
        if (lane == 0) s_red[k][warp] = r;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int k = threadIdx.x;
        const bool is_min = (k < 3) || (k == 6);
        uint32_t r = s_red[k][0];
        for (int w = 1; w < kAabbThreads / 32; w++) r = is_min ? min(r, s_red[k][w]) : max(r, s_red[k][w]);
        if (k < 6 || radii) {
            if (is_min) atomicMin(&out[k], r); else atomicMax(&out[k], r);
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// Cell assignment + Morton key.  ijk = floor((p - bottom) * inv_cell) like TreeNSearch.cpp:713-715, but in fp64 and clamped.
template <typename Key>
__global__ void __launch_bounds__(256) keygen_kernel(const float* __restrict__ pts, int n, int stride, GridParams g, Key* __restrict__ keys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = pts + (int64_t)i * stride;
    const float x = p[0], y = p[1], z = p[2];
    int cx = __double2int_rd(((double)x - g.bottom[0]) * g.inv_cell);
    int cy = __double2int_rd(((double)y - g.bottom[1]) * g.inv_cell);
    int cz = __double2int_rd(((double)z - g.bottom[2]) * g.inv_cell);
    cx = min(max(cx, 0), g.max_coord);
    cy = min(max(cy, 0), g.max_coord);
    cz = min(max(cz, 0), g.max_coord);
    keys[i] = Morton<Key>::encode((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
}

// ----------------------------------------------------------------------------------------------------------------------
// Reorder: sorted[i] = (x, y, z, bits(idx)) of the point that the sort put at position i; r2[i] = r*r (float, like
// TreeNSearch.cpp:2352) in variable radius mode.  ids[] (optional) replaces the set-local index by a caller supplied id.
__global__ void __launch_bounds__(256) reorder_kernel(const float* __restrict__ pts, int stride, const float* __restrict__ radii, const uint32_t* __restrict__ order,
                                                      int n, float4* __restrict__ sorted, float* __restrict__ sorted_r2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = order[i];
    const float* p = pts + (int64_t)src * stride;
    sorted[i] = make_float4(p[0], p[1], p[2], __uint_as_float(src));
    if (radii) {
        const float r = radii[src];
        sorted_r2[i] = __fmul_rn(r, r);
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// Cell start/end compaction over the sorted keys + open addressing hash (cell key -> compact cell id).
constexpr int kCellThreads = 256;
constexpr int kCellItems = 4;
constexpr int kCellTile = kCellThreads * kCellItems;

__device__ __forceinline__ uint32_t cas_key(uint32_t* a, uint32_t expected, uint32_t v) { return atomicCAS(a, expected, v); }
__device__ __forceinline__ uint64_t cas_key(uint64_t* a, uint64_t expected, uint64_t v)
{
    return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long*>(a), (unsigned long long)expected, (unsigned long long)v);
}

template <typename Key>
__global__ void __launch_bounds__(kCellThreads) count_heads_kernel(const Key* __restrict__ keys, int n, uint32_t* __restrict__ tile_heads)
{
    __shared__ uint32_t warp_sums[8];
    const int base = blockIdx.x * kCellTile + threadIdx.x * kCellItems;
    uint32_t c = 0;
    Key prev = (base > 0 && base - 1 < n) ? keys[base - 1] : (Key)0;
#pragma unroll
    for (int i = 0; i < kCellItems; i++) {
        const int idx = base + i;
        if (idx < n) {
            const Key k = keys[idx];
            c += (idx == 0 || k != prev) ? 1u : 0u;
            prev = k;
        }
    }
    uint32_t total;
    block_exclusive_scan_256(c, warp_sums, total);
    if (threadIdx.x == 0) tile_heads[blockIdx.x] = total;
}

template <typename Key>
__global__ void __launch_bounds__(kCellThreads) emit_cells_kernel(const Key* __restrict__ keys, int n, const uint32_t* __restrict__ tile_base,
                                                                  Key* __restrict__ cell_key, uint32_t* __restrict__ cell_start)
{
    __shared__ uint32_t warp_sums[8];
    const int base = blockIdx.x * kCellTile + threadIdx.x * kCellItems;
    Key k[kCellItems];
    bool head[kCellItems];
    uint32_t c = 0;
    Key prev = (base > 0 && base - 1 < n) ? keys[base - 1] : (Key)0;
#pragma unroll
    for (int i = 0; i < kCellItems; i++) {
        const int idx = base + i;
        head[i] = false;
        k[i] = (Key)0;
        if (idx < n) {
            k[i] = keys[idx];
            head[i] = (idx == 0 || k[i] != prev);
            prev = k[i];
            c += head[i] ? 1u : 0u;
        }
    }
    uint32_t total;
    uint32_t cid = block_exclusive_scan_256(c, warp_sums, total) + tile_base[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kCellItems; i++) {
        if (head[i]) {
            cell_key[cid] = k[i];
            cell_start[cid] = (uint32_t)(base + i);
            cid++;
        }
    }
    // sentinel: cell_start[n_cells] = n
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kCellThreads - 1) cell_start[cid] = (uint32_t)n;
}

// hash insert of every occupied cell: key -> [start, end)
template <typename Key>
__global__ void __launch_bounds__(256) build_hash_kernel(const Key* __restrict__ cell_key, const uint32_t* __restrict__ cell_start, int n_cells,
                                                         typename HashSlot<Key>::Raw* __restrict__ table, int hash_log2)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const Key k = cell_key[c];
    const uint32_t s = cell_start[c], e = cell_start[c + 1];
    const uint32_t hmask = (1u << hash_log2) - 1u;
    uint32_t slot = Morton<Key>::hash(k) >> (32 - hash_log2);
    while (!HashSlot<Key>::try_insert(table, slot, k, s, e)) slot = (slot + 1) & hmask;
}

// dense Morton-indexed cell table {start, end}: fill for the current cells / clear the entries of the previous run
template <typename Key>
__global__ void __launch_bounds__(256) dense_table_kernel(const Key* __restrict__ cell_key, const uint32_t* __restrict__ cell_start, int n_cells,
                                                          uint2* __restrict__ table, int fill)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    table[cell_key[c]] = fill ? make_uint2(cell_start[c], cell_start[c + 1]) : make_uint2(0u, 0u);
}

// ----------------------------------------------------------------------------------------------------------------------
// Bucket build: when the cell table (one counter per cell of the whole grid) is small next to the point count, the sort of
// (cell key, index) pairs collapses into ONE counting pass over the full key -- histogram, exclusive scan, scatter -- instead
// of ceil(key_bits / 8) LSD radix passes, and the scatter writes the reordered float4 records directly (no separate gather).
//   keygen_count_kernel   key of every point + population of its cell (fire-and-forget L2 atomics on an L2 resident table)
//   exclusive_scan_u32    first[k] = number of points with a smaller key, for every k in [0, 2^key_bits]
//   bucket_scatter_kernel sorted[cursor[key]++] = (x, y, z, bits(index)); cursor starts as a copy of first.  (Measured: taking the rank
//                         from a value-returning atomic in the first kernel instead is slower -- the scatter is bound by its random
//                         16-byte writes, not by its atomics, and returning atomics cost 35 us more in keygen.)
//   table_count/emit      compact list of the occupied cells (the query's task list) + the cell kernel's {start, end} table
// The order of the points INSIDE a cell is the arrival order of the atomics (not the input order as with the stable radix
// sort); neighbour SETS do not depend on it.  prepare_zsort() and TNSB_OPT_BUILD = 1 use the radix path.
template <typename Key>
__global__ void __launch_bounds__(256) keygen_count_kernel(const float* __restrict__ pts, int n, int stride, GridParams g, Key* __restrict__ keys,
                                                           uint32_t* __restrict__ population)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = pts + (int64_t)i * stride;
    const float x = p[0], y = p[1], z = p[2];
    int cx = __double2int_rd(((double)x - g.bottom[0]) * g.inv_cell);
    int cy = __double2int_rd(((double)y - g.bottom[1]) * g.inv_cell);
    int cz = __double2int_rd(((double)z - g.bottom[2]) * g.inv_cell);
    cx = min(max(cx, 0), g.max_coord);
    cy = min(max(cy, 0), g.max_coord);
    cz = min(max(cz, 0), g.max_coord);
    const Key k = Morton<Key>::encode((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
    keys[i] = k;
    atomicAdd(population + k, 1u);
}

// the same pass for the brick query's grid (query_brick.cuh): half-radius cells, linear row keys  key = (z * ny + y) * nx + x.
// The cell arithmetic must stay identical to brick_cell() in query_brick.cuh (fp64 subtract, multiply, floor, clamp).
__global__ void __launch_bounds__(256) brick_keygen_count_kernel(const float* __restrict__ pts, int n, int stride, BrickGrid g, uint32_t* __restrict__ keys,
                                                                 uint32_t* __restrict__ population)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = pts + (int64_t)i * stride;
    const float x = p[0], y = p[1], z = p[2];
    int cx = __double2int_rd(((double)x - g.bottom[0]) * g.inv_cell);
    int cy = __double2int_rd(((double)y - g.bottom[1]) * g.inv_cell);
    int cz = __double2int_rd(((double)z - g.bottom[2]) * g.inv_cell);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    cz = min(max(cz, 0), g.nz - 1);
    const uint32_t k = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
    keys[i] = k;
    atomicAdd(population + k, 1u);
}

constexpr int kScatterItems = 4;     // points per thread: the pass is bound by the latency of its dependent chain key -> atomic -> store, so every
                                     // thread keeps four chains in flight (ncu: 90 % of the stall samples were long_scoreboard with one chain per thread)

template <typename Key>
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const float* __restrict__ pts, int stride, const float* __restrict__ radii, const Key* __restrict__ keys,
                                                             int n, uint32_t* __restrict__ cursor, float4* __restrict__ sorted, float* __restrict__ sorted_r2,
                                                             Key key_lo, Key key_hi)
{
    // [key_lo, key_hi): destination window of this launch.  Random 16-byte writes over a 160 MB output miss L2 and make DRAM
    // read-modify-write half sectors (ncu: 374 MB read + 287 MB written for 160 MB of records); with a window that fits L2 both
    // halves of a sector meet there.  Measured at 10M points: 1 window 0.319 ms, 2 windows 0.271 ms, 4: 0.286 ms, 8: 0.427 ms.
    const int base = blockIdx.x * (256 * kScatterItems) + threadIdx.x;
    Key k[kScatterItems];
    bool in[kScatterItems];
#pragma unroll
    for (int u = 0; u < kScatterItems; u++) {
        const int i = base + u * 256;
        in[u] = false;
        if (i < n) {
            k[u] = keys[i];
            in[u] = k[u] >= key_lo && k[u] < key_hi;
        }
    }
    float x[kScatterItems], y[kScatterItems], z[kScatterItems], r[kScatterItems];
    uint32_t pos[kScatterItems];
#pragma unroll
    for (int u = 0; u < kScatterItems; u++) {
        if (in[u]) {
            const int i = base + u * 256;
            const float* p = pts + (int64_t)i * stride;
            x[u] = p[0]; y[u] = p[1]; z[u] = p[2];
            if (radii) r[u] = radii[i];
            pos[u] = atomicAdd(cursor + k[u], 1u);
        }
    }
#pragma unroll
    for (int u = 0; u < kScatterItems; u++) {
        if (in[u]) {
            const int i = base + u * 256;
            sorted[pos[u]] = make_float4(x[u], y[u], z[u], __uint_as_float((uint32_t)i));
            if (radii) sorted_r2[pos[u]] = __fmul_rn(r[u], r[u]);
        }
    }
}

// Speculative reuse of the previous run's brick grid (no host round trip for the world box): one thread checks the fresh
// box / radius range in `red` (ordered-uint encoding, see aabb_kernel) against the grid the build is about to use and raises *flag
// when a point would fall outside it, or when the largest radius no longer matches the cell size.  The host reads the flag together
// with the query counters at the end of the run and repeats the run with a fresh grid if it is set.
__global__ void box_check_kernel(const uint32_t* __restrict__ red, BrickGrid g, float r_max_grid, int variable_radius, int* __restrict__ flag)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bool bad = false;
    const int dims[3] = { g.nx, g.ny, g.nz };
    for (int d = 0; d < 3; d++) {
        const float lo = ordered_to_float(red[d]), hi = ordered_to_float(red[3 + d]);
        if (!(lo == lo) || !(hi == hi) || isinf(lo) || isinf(hi)) bad = true;
        const double clo = floor(((double)lo - g.bottom[d]) * g.inv_cell), chi = floor(((double)hi - g.bottom[d]) * g.inv_cell);
        if (clo < 0.0 || chi > (double)(dims[d] - 1)) bad = true;
    }
    if (variable_radius) {
        const float r_lo = ordered_to_float(red[6]), r_hi = ordered_to_float(red[7]);
        if (!(r_lo > 0.0f) || !(r_hi <= r_max_grid) || r_hi < 0.8f * r_max_grid) bad = true;
    }
    if (bad) *flag = 1;
}

// occupied cells of the prefix table first[0 .. n_keys]: count per tile, then (after a scan of the tile counts) emit
// cell_key / cell_start in key order; optionally fills the cell kernel's dense {start, end} table for every key
__global__ void __launch_bounds__(kCellThreads) table_count_kernel(const uint32_t* __restrict__ first, int64_t n_keys, uint32_t* __restrict__ tile_cells)
{
    __shared__ uint32_t warp_sums[8];
    const int64_t base = (int64_t)blockIdx.x * kCellTile + (int64_t)threadIdx.x * kCellItems;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < kCellItems; i++) {
        const int64_t k = base + i;
        if (k < n_keys) c += first[k + 1] > first[k] ? 1u : 0u;
    }
    uint32_t total;
    block_exclusive_scan_256(c, warp_sums, total);
    if (threadIdx.x == 0) tile_cells[blockIdx.x] = total;
}

template <typename Key>
__global__ void __launch_bounds__(kCellThreads) table_emit_kernel(const uint32_t* __restrict__ first, int64_t n_keys, const uint32_t* __restrict__ tile_base,
                                                                  Key* __restrict__ cell_key, uint32_t* __restrict__ cell_start, uint2* __restrict__ dense)
{
    __shared__ uint32_t warp_sums[8];
    const int64_t base = (int64_t)blockIdx.x * kCellTile + (int64_t)threadIdx.x * kCellItems;
    uint32_t lo[kCellItems + 1];
#pragma unroll
    for (int i = 0; i <= kCellItems; i++) lo[i] = (base + i <= n_keys) ? first[base + i] : 0u;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < kCellItems; i++)
        if (base + i < n_keys) {
            c += lo[i + 1] > lo[i] ? 1u : 0u;
            if (dense) dense[base + i] = make_uint2(lo[i], lo[i + 1]);
        }
    uint32_t total;
    uint32_t cid = block_exclusive_scan_256(c, warp_sums, total) + tile_base[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kCellItems; i++) {
        if (base + i < n_keys && lo[i + 1] > lo[i]) {
            cell_key[cid] = (Key)(base + i);
            cell_start[cid] = lo[i];
            cid++;
        }
    }
    // sentinel: cell_start[n_cells] = n (= first[n_keys])
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kCellThreads - 1) cell_start[cid] = first[n_keys];
}

// prefix cell table, step 1: population of every occupied cell at table[key] (the table was zeroed); an exclusive scan then turns
// it into first[key] = number of points with a smaller key, valid for EVERY key in [0, 2^key_bits] (empty cells included)
template <typename Key>
__global__ void __launch_bounds__(256) cell_population_kernel(const Key* __restrict__ cell_key, const uint32_t* __restrict__ cell_start, int n_cells,
                                                              uint32_t* __restrict__ table)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    table[cell_key[c]] = cell_start[c + 1] - cell_start[c];
}

// prepare_zsort() on top of a resident grid: Morton key (cell = r grid, libmorton order) of every SORTED record, written at the
// record's original index, so that a stable radix sort of (key, index) yields the same order as a build from the raw arrays
template <typename Key>
__global__ void __launch_bounds__(256) zsort_keys_kernel(const float4* __restrict__ sorted, int n, GridParams g, Key* __restrict__ keys_by_index)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = sorted[i];
    int cx = __double2int_rd(((double)v.x - g.bottom[0]) * g.inv_cell);
    int cy = __double2int_rd(((double)v.y - g.bottom[1]) * g.inv_cell);
    int cz = __double2int_rd(((double)v.z - g.bottom[2]) * g.inv_cell);
    cx = min(max(cx, 0), g.max_coord);
    cy = min(max(cy, 0), g.max_coord);
    cz = min(max(cz, 0), g.max_coord);
    keys_by_index[__float_as_int(v.w)] = Morton<Key>::encode((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
}

// fused multi-array gather of apply_zsort on the device: dst_k[row] = src_k[new_to_old[row]] for up to 8 arrays of 4-byte words
constexpr int kMaxZsortArrays = 8;
struct ZsortArrays {
    const uint32_t* src[kMaxZsortArrays];
    uint32_t* dst[kMaxZsortArrays];
    int row_words[kMaxZsortArrays];
    int n_arrays;
};
__global__ void __launch_bounds__(256) gather_arrays_kernel(ZsortArrays a, const int32_t* __restrict__ new_to_old, int n)
{
    const int k = blockIdx.y;
    const int w = a.row_words[k];
    const int64_t total = (int64_t)n * w;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(t / w), c = (int)(t - (int64_t)row * w);
        a.dst[k][t] = a.src[k][(int64_t)new_to_old[row] * w + c];
    }
}

// gather used by tnsb_apply_zsort_device_f32
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, const int32_t* __restrict__ new_to_old,
                                                          int n, int stride)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * stride) return;
    const int row = (int)(t / stride), c = (int)(t % stride);
    dst[t] = src[(int64_t)new_to_old[row] * stride + c];
}

}  // namespace tnsb
