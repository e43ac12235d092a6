// scan.cuh -- device-wide exclusive prefix sum over uint32 (three launches: tile sums, scan of tile sums, apply).
// Used for the radix sort's digit/tile matrix and for cell-head compaction; inputs are small (<= a few MB), so the
// straightforward reduce-then-scan organisation is used instead of a single-pass chained scan.
#pragma once
#include "common.cuh"

namespace tnsb {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;   // 2048

// exclusive scan of one value per thread across a 256-thread block; returns the exclusive prefix, total in `total`
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* warp_sums /*[8] shared*/, uint32_t& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    total = tot;
    __syncthreads();     // warp_sums may be reused by the caller
    return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ tile_sums, int64_t n)
{
    __shared__ uint32_t warp_sums[8];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        const int64_t idx = base + threadIdx.x + (int64_t)i * kScanThreads;
        if (idx < n) s += in[idx];
    }
    uint32_t total;
    block_exclusive_scan_256(s, warp_sums, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sums in place; writes the grand total to *total_out (if not null)
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(uint32_t* __restrict__ tile_sums, int n_tiles, uint32_t* __restrict__ total_out)
{
    __shared__ uint32_t warp_sums[8];
    uint32_t running = 0;
    for (int base = 0; base < n_tiles; base += kScanThreads) {
        const int idx = base + threadIdx.x;
        const uint32_t v = idx < n_tiles ? tile_sums[idx] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan_256(v, warp_sums, total);
        if (idx < n_tiles) tile_sums[idx] = running + ex;
        running += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = running;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                  const uint32_t* __restrict__ tile_sums, int64_t n)
{
    __shared__ uint32_t warp_sums[8];
    // blocked arrangement: thread t owns kScanItems consecutive values
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan_256(s, warp_sums, total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

// temp must hold ceil(n / kScanTile) uint32.  In-place (out == in) is allowed.  total_out (device pointer) may be null.
inline int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* temp, uint32_t* total_out, cudaStream_t stream)
{
    if (n <= 0) {
        if (total_out) cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream);
        return 0;
    }
    const int n_tiles = (int)ceil_div64(n, kScanTile);
    scan_tile_sums_kernel<<<n_tiles, kScanThreads, 0, stream>>>(in, temp, n);
    scan_partials_kernel<<<1, kScanThreads, 0, stream>>>(temp, n_tiles, total_out);
    scan_apply_kernel<<<n_tiles, kScanThreads, 0, stream>>>(in, out, temp, n);
    return 3;
}

inline int64_t exclusive_scan_temp_elems(int64_t n) { return ceil_div64(n > 0 ? n : 1, kScanTile) + 1; }

}  // namespace tnsb
