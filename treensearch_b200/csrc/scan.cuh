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

// out2 (optional): a second copy of the result limited to its first n2 entries (the bucket build's scatter cursors = the prefix table
// without its last entry), written here instead of by a separate device-to-device copy.  A thread owns kScanItems = 8 consecutive values:
// two 128-bit loads / stores when the arrays are 16-byte aligned (a scalar 4-byte store per value makes eight partial writes per sector).
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                  const uint32_t* __restrict__ tile_sums, int64_t n, uint32_t* __restrict__ out2, int64_t n2, int vec_ok)
{
    __shared__ uint32_t warp_sums[8];
    // blocked arrangement: thread t owns kScanItems consecutive values
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    const bool full = vec_ok && base + kScanItems <= n;
    if (full) {
        const uint4 a = reinterpret_cast<const uint4*>(in + base)[0], b = reinterpret_cast<const uint4*>(in + base)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++) v[i] = (base + i < n) ? in[base + i] : 0u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) s += v[i];
    uint32_t total;
    uint32_t ex = block_exclusive_scan_256(s, warp_sums, total) + tile_sums[blockIdx.x];
    uint32_t r[kScanItems];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        r[i] = ex;
        ex += v[i];
    }
    if (full) {
        const uint4 a = make_uint4(r[0], r[1], r[2], r[3]), b = make_uint4(r[4], r[5], r[6], r[7]);
        reinterpret_cast<uint4*>(out + base)[0] = a;
        reinterpret_cast<uint4*>(out + base)[1] = b;
        if (out2 && base + kScanItems <= n2) {
            reinterpret_cast<uint4*>(out2 + base)[0] = a;
            reinterpret_cast<uint4*>(out2 + base)[1] = b;
        } else if (out2) {
#pragma unroll
            for (int i = 0; i < kScanItems; i++)
                if (base + i < n2) out2[base + i] = r[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++) {
            if (base + i < n) out[base + i] = r[i];
            if (out2 && base + i < n2) out2[base + i] = r[i];
        }
    }
}

// temp must hold ceil(n / kScanTile) uint32.  In-place (out == in) is allowed.  total_out (device pointer) may be null.
inline int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* temp, uint32_t* total_out, cudaStream_t stream, uint32_t* out2 = nullptr,
                              int64_t n2 = 0)
{
    if (n <= 0) {
        if (total_out) cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream);
        return 0;
    }
    const int n_tiles = (int)ceil_div64(n, kScanTile);
    scan_tile_sums_kernel<<<n_tiles, kScanThreads, 0, stream>>>(in, temp, n);
    scan_partials_kernel<<<1, kScanThreads, 0, stream>>>(temp, n_tiles, total_out);
    const int vec_ok = (((uintptr_t)in | (uintptr_t)out | (uintptr_t)out2) & 15u) == 0 ? 1 : 0;
    scan_apply_kernel<<<n_tiles, kScanThreads, 0, stream>>>(in, out, temp, n, out2, n2, vec_ok);
    return 3;
}

inline int64_t exclusive_scan_temp_elems(int64_t n) { return ceil_div64(n > 0 ? n : 1, kScanTile) + 1; }

}  // namespace tnsb
