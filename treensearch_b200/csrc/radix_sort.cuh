// radix_sort.cuh -- stable LSD radix sort of (Morton key, point index) pairs, 8 bits per pass, hand written for sm_100a.
//
// Replaces the role of Taskflow's parallel pdqsort in the reference (TreeNSearch.cpp:2624, :2702-2705) and, more
// importantly, turns the reference's "runs of consecutive points in one cell" (TreeNSearch.cpp:646-749) into true cells.
//
// One pass = three steps:
//   1. radix_hist_kernel    per-tile digit histogram          -> hist[digit][tile]   (digit-major)
//   2. exclusive_scan_u32   over the flattened matrix         -> global base of (digit, tile)
//   3. radix_scatter_kernel stable in-tile ranking (warp match_any multisplit), staging of the tile in digit order in
//                           shared memory, then coalesced runs to global memory.
// Only as many passes as the key has significant bits are run (3 passes for a 7-bit-per-axis grid).
//
// Algorithmic traffic per pass and pair: read key (4|8 B) + value 4 B, write the same; plus one extra key read for (1).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace tnsb {

constexpr int kRadixBits = 8;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;

template <typename Key> struct SortTile;
template <> struct SortTile<uint32_t> { static constexpr int kItems = 16; };
template <> struct SortTile<uint64_t> { static constexpr int kItems = 8; };

template <typename Key>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const Key* __restrict__ keys, int n, int n_tiles, int shift, uint32_t mask,
                                                                  uint32_t* __restrict__ hist /*[kRadixBins][n_tiles]*/)
{
    constexpr int kTile = kSortThreads * SortTile<Key>::kItems;
    __shared__ uint32_t s_hist[kRadixBins];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kTile;
#pragma unroll
    for (int i = 0; i < SortTile<Key>::kItems; i++) {
        const int idx = base + i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&s_hist[(uint32_t)(keys[idx] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_hist[threadIdx.x];
}

template <typename Key>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const Key* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                     Key* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                     int n, int n_tiles, int shift, uint32_t mask,
                                                                     const uint32_t* __restrict__ scanned /*[kRadixBins][n_tiles]*/)
{
    constexpr int kItems = SortTile<Key>::kItems;
    constexpr int kTile = kSortThreads * kItems;
    constexpr int kPerWarp = kItems * 32;

    __shared__ uint32_t s_whist[kSortWarps][kRadixBins];   // per-warp digit counts, later exclusive bases over warps
    __shared__ uint32_t s_tile_start[kRadixBins];          // first staged slot of each digit
    __shared__ uint32_t s_gbase[kRadixBins];               // global index of staged slot p with digit d is s_gbase[d] + p
    __shared__ uint32_t s_warp_sums[8];
    __shared__ Key s_keys[kTile];
    __shared__ uint32_t s_vals[kTile];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_base = blockIdx.x * kTile;
    const int tile_n = min(kTile, n - tile_base);

#pragma unroll
    for (int w = 0; w < kSortWarps; w++) s_whist[w][threadIdx.x] = 0;
    __syncthreads();

    // ---- stable ranking inside the warp's contiguous slice (kItems rounds of 32 consecutive keys)
    Key key[kItems];
    uint32_t val[kItems];
    uint16_t rank[kItems];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const int loc = warp * kPerWarp + r * 32 + lane;
        const int idx = tile_base + loc;
        const bool valid = loc < tile_n;
        key[r] = valid ? keys_in[idx] : (Key)0;
        val[r] = valid ? (vals_in ? vals_in[idx] : (uint32_t)idx) : 0u;
        const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
        const unsigned peers = __match_any_sync(kFull, valid ? d : (uint32_t)kRadixBins);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = s_whist[warp][d];
            s_whist[warp][d] = pre + __popc(peers);
        }
        pre = __shfl_sync(kFull, pre, leader);
        rank[r] = (uint16_t)(pre + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive prefix over warps, tile total, exclusive prefix over digits
    {
        const int d = threadIdx.x;
        uint32_t running = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const uint32_t c = s_whist[w][d];
            s_whist[w][d] = running;
            running += c;
        }
        uint32_t total;
        const uint32_t start = block_exclusive_scan_256(running, s_warp_sums, total);
        s_tile_start[d] = start;
        s_gbase[d] = scanned[(size_t)d * n_tiles + blockIdx.x] - start;
    }
    __syncthreads();

    // ---- stage the tile in digit order
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const int loc = warp * kPerWarp + r * 32 + lane;
        if (loc < tile_n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            const uint32_t p = s_tile_start[d] + s_whist[warp][d] + rank[r];
            s_keys[p] = key[r];
            s_vals[p] = val[r];
        }
    }
    __syncthreads();

    // ---- coalesced runs to global memory
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const int p = i * kSortThreads + threadIdx.x;
        if (p < tile_n) {
            const Key k = s_keys[p];
            const uint32_t d = (uint32_t)(k >> shift) & mask;
            const uint32_t g = s_gbase[d] + (uint32_t)p;
            keys_out[g] = k;
            vals_out[g] = s_vals[p];
        }
    }
}

template <typename Key>
inline int64_t radix_sort_temp_elems(int n)
{
    const int kTile = kSortThreads * SortTile<Key>::kItems;
    const int64_t n_tiles = ceil_div(n > 0 ? n : 1, kTile);
    const int64_t m = n_tiles * kRadixBins;
    return m + exclusive_scan_temp_elems(m);
}

// Sorts n pairs by bits [0, key_bits) of the key.  Values of the first pass are the identity permutation (vals[0] is never read).
// keys[2] / vals[2] are ping-pong buffers; returns the index (0/1) of the buffers that hold the result.  *launches is incremented.
template <typename Key>
inline int radix_sort_pairs(Key* keys[2], uint32_t* vals[2], int n, int key_bits, uint32_t* temp, cudaStream_t stream, int* launches, int* passes_out)
{
    int sel = 0, passes = 0;
    if (n > 0) {
        constexpr int kTile = kSortThreads * SortTile<Key>::kItems;
        const int n_tiles = ceil_div(n, kTile);
        const int64_t m = (int64_t)n_tiles * kRadixBins;
        uint32_t* hist = temp;
        uint32_t* scan_temp = temp + m;
        for (int shift = 0; shift < key_bits || passes == 0; shift += kRadixBits) {
            const int bits = min(kRadixBits, max(key_bits - shift, 1));
            const uint32_t mask = (1u << bits) - 1u;
            radix_hist_kernel<Key><<<n_tiles, kSortThreads, 0, stream>>>(keys[sel], n, n_tiles, shift, mask, hist);
            *launches += 1 + exclusive_scan_u32(hist, hist, m, scan_temp, nullptr, stream);
            radix_scatter_kernel<Key><<<n_tiles, kSortThreads, 0, stream>>>(keys[sel], passes == 0 ? nullptr : vals[sel], keys[sel ^ 1], vals[sel ^ 1],
                                                                           n, n_tiles, shift, mask, hist);
            *launches += 1;
            sel ^= 1;
            passes++;
        }
    }
    if (passes_out) *passes_out = passes;
    return sel;
}

}  // namespace tnsb
