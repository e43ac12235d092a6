// shard.cuh -- device helpers of the multi-GPU path: Z-slab decomposition of a point cloud across the GPUs of one box.
// The reference is single-process shared memory (no counterpart file); the decomposition follows SURVEY.md §8e:
//   every GPU owns the points of one slab along an axis, cut so that the slabs hold equal point counts, and additionally
//   receives a read-only halo of width >= r_max on both sides, which is all a fixed-radius query needs.
// Kernels: coordinate histogram along the axis (for the balanced cuts) and a two-pass bucket partition of the local
// chunk into  [owned -> GPU 0 .. G-1 | halo -> GPU 0 .. G-1]  records (x, y, z, bits(global id)), ready for all_to_all.
#pragma once
#include "common.cuh"

namespace tnsb {

constexpr int kShardThreads = 256;
constexpr int kMaxParts = 64;

__global__ void __launch_bounds__(kShardThreads) axis_histogram_kernel(const float* __restrict__ pts, int n, int stride, int axis, float lo, float inv_bin,
                                                                       int n_bins, uint32_t* __restrict__ hist)
{
    extern __shared__ uint32_t s_hist[];
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = pts[(size_t)i * stride + axis];
        int b = (int)((v - lo) * inv_bin);
        b = min(max(b, 0), n_bins - 1);
        atomicAdd(&s_hist[b], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x)
        if (s_hist[b]) atomicAdd(&hist[b], s_hist[b]);
}

struct SlabCuts {
    float cut[kMaxParts + 1];   // part g owns [cut[g], cut[g+1]); cut[0] = -inf, cut[n_parts] = +inf
    int n_parts;
    float halo;
};

__device__ __forceinline__ int slab_of(const SlabCuts& c, float v)
{
    int g = 0;
    for (int k = 1; k < c.n_parts; k++) g += (v >= c.cut[k]) ? 1 : 0;
    return g;
}

// pass 1: counts[g] = points owned by g, counts[n_parts + g] = halo copies for g
__global__ void __launch_bounds__(kShardThreads) slab_count_kernel(const float* __restrict__ pts, int n, int stride, int axis, SlabCuts c,
                                                                   unsigned long long* __restrict__ counts)
{
    __shared__ uint32_t s_cnt[2 * kMaxParts];
    for (int b = threadIdx.x; b < 2 * c.n_parts; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = pts[(size_t)i * stride + axis];
        const int g = slab_of(c, v);
        atomicAdd(&s_cnt[g], 1u);
        const int g_lo = slab_of(c, v - c.halo), g_hi = slab_of(c, v + c.halo);
        for (int h = g_lo; h <= g_hi; h++)
            if (h != g) atomicAdd(&s_cnt[c.n_parts + h], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * c.n_parts; b += blockDim.x)
        if (s_cnt[b]) atomicAdd(&counts[b], (unsigned long long)s_cnt[b]);
}

// pass 2: scatter records; cursors[b] starts at the bucket's first slot (exclusive prefix of counts) and is advanced with
// one atomic per (block, bucket) so that every block writes contiguous runs
__global__ void __launch_bounds__(kShardThreads) slab_scatter_kernel(const float* __restrict__ pts, int n, int stride, int axis, int id_base, SlabCuts c,
                                                                     unsigned long long* __restrict__ cursors, float4* __restrict__ out)
{
    __shared__ uint32_t s_cnt[2 * kMaxParts];
    __shared__ unsigned long long s_base[2 * kMaxParts];
    const int per_block = (n + gridDim.x - 1) / gridDim.x;
    const int begin = blockIdx.x * per_block, end = min(n, begin + per_block);
    for (int b = threadIdx.x; b < 2 * c.n_parts; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const float v = pts[(size_t)i * stride + axis];
        const int g = slab_of(c, v);
        atomicAdd(&s_cnt[g], 1u);
        const int g_lo = slab_of(c, v - c.halo), g_hi = slab_of(c, v + c.halo);
        for (int h = g_lo; h <= g_hi; h++)
            if (h != g) atomicAdd(&s_cnt[c.n_parts + h], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * c.n_parts; b += blockDim.x) {
        s_base[b] = s_cnt[b] ? atomicAdd(&cursors[b], (unsigned long long)s_cnt[b]) : 0ull;
        s_cnt[b] = 0;
    }
    __syncthreads();
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const float* p = pts + (size_t)i * stride;
        const float4 rec = make_float4(p[0], p[1], p[2], __int_as_float(id_base + i));
        const float v = p[axis];
        const int g = slab_of(c, v);
        out[s_base[g] + atomicAdd(&s_cnt[g], 1u)] = rec;
        const int g_lo = slab_of(c, v - c.halo), g_hi = slab_of(c, v + c.halo);
        for (int h = g_lo; h <= g_hi; h++)
            if (h != g) out[s_base[c.n_parts + h] + atomicAdd(&s_cnt[c.n_parts + h], 1u)] = rec;
    }
}


// ---- one-sided exchange: every rank owns a RECEIVE WINDOW in its HBM, mapped into every peer process through CUDA IPC; the
// partition kernel pushes each record straight into its owner's window (and into the windows of the ranks that need it as
// halo) over NVLink -- partition and exchange are ONE kernel, there is no count pass, no count exchange and no all-to-all.
// Window layout: [header 256 B: u64 owned count, u64 halo count | cap_owned float4 records | cap_halo float4 records].
// One system-scope atomic per (block, destination bucket) reserves a contiguous run in the destination window; records
// beyond a window's capacity are dropped while the counter keeps counting; the pusher raises `flag` to 2 (and the owner's
// tnsb_shard_collect reports TNSB_ERR_LIMIT).
constexpr int kWindowHeaderBytes = 256;

struct PushWindows {
    char* base[kMaxParts];      // window of every rank (this rank's own entry is plain local memory)
    long long cap_owned, cap_halo;
};

__global__ void __launch_bounds__(kShardThreads) slab_push_kernel(const float* __restrict__ pts, int n, int stride, int axis, int id_base, SlabCuts c, PushWindows w,
                                                                  int* __restrict__ flag)
{
    __shared__ uint32_t s_cnt[2 * kMaxParts];
    __shared__ unsigned long long s_base[2 * kMaxParts];
    const int per_block = (n + gridDim.x - 1) / gridDim.x;
    const int begin = blockIdx.x * per_block, end = min(n, begin + per_block);
    for (int b = threadIdx.x; b < 2 * c.n_parts; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const float v = pts[(size_t)i * stride + axis];
        const int g = slab_of(c, v);
        atomicAdd(&s_cnt[g], 1u);
        const int g_lo = slab_of(c, v - c.halo), g_hi = slab_of(c, v + c.halo);
        for (int h = g_lo; h <= g_hi; h++)
            if (h != g) atomicAdd(&s_cnt[c.n_parts + h], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * c.n_parts; b += blockDim.x) {
        const int g = b < c.n_parts ? b : b - c.n_parts;
        unsigned long long* counter = reinterpret_cast<unsigned long long*>(w.base[g]) + (b < c.n_parts ? 0 : 1);
        s_base[b] = s_cnt[b] ? atomicAdd_system(counter, (unsigned long long)s_cnt[b]) : 0ull;
        // the reservation does not fit the destination window: tell the host (the flag rides on the barrier all_reduce, so that
        // EVERY rank repeats the step with larger windows)
        if (s_cnt[b] && flag && (long long)(s_base[b] + s_cnt[b]) > (b < c.n_parts ? w.cap_owned : w.cap_halo)) atomicMax(flag, 2);
        s_cnt[b] = 0;
    }
    __syncthreads();
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const float* p = pts + (size_t)i * stride;
        const float4 rec = make_float4(p[0], p[1], p[2], __int_as_float(id_base + i));
        const float v = p[axis];
        const int g = slab_of(c, v);
        {
            const unsigned long long pos = s_base[g] + atomicAdd(&s_cnt[g], 1u);
            if ((long long)pos < w.cap_owned) reinterpret_cast<float4*>(w.base[g] + kWindowHeaderBytes)[pos] = rec;
        }
        const int g_lo = slab_of(c, v - c.halo), g_hi = slab_of(c, v + c.halo);
        for (int h = g_lo; h <= g_hi; h++) {
            if (h == g) continue;
            const unsigned long long pos = s_base[c.n_parts + h] + atomicAdd(&s_cnt[c.n_parts + h], 1u);
            if ((long long)pos < w.cap_halo) reinterpret_cast<float4*>(w.base[h] + kWindowHeaderBytes)[w.cap_owned + pos] = rec;
        }
    }
}

}  // namespace tnsb
