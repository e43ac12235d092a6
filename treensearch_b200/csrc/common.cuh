// common.cuh -- shared device helpers of the B200 neighbour-search engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tnsb {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// grid geometry shared by key generation and the query (device copy lives in constant-like kernel arguments)
struct GridParams {
    double bottom[3];   // world box origin (double: cell coordinates are computed in fp64 so that the stencil argument holds
    double inv_cell;    //                   for any grid resolution, see DESIGN.md "candidate completeness")
    int    bits;        // Morton bits per dimension (cells per dimension = 1 << bits)
    int    max_coord;   // (1 << bits) - 1
};

// grid of the brick query (query_brick.cuh): half-radius cells, nx x ny x nz of them, linear row keys (z * ny + y) * nx + x
struct BrickGrid {
    double bottom[3];
    double inv_cell;
    int nx, ny, nz;
};

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- Morton codes, libmorton bit order (x -> bit 0, y -> bit 1, z -> bit 2; reference: extern/libmorton/morton_BMI.h:40-52)
__host__ __device__ __forceinline__ uint32_t expand_bits_10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t compact_bits_10(uint32_t v)
{
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030c30c3u;
    v = (v | (v >> 4)) & 0x0300f00fu;
    v = (v | (v >> 8)) & 0x030000ffu;
    v = (v | (v >> 16)) & 0x3ffu;
    return v;
}
__host__ __device__ __forceinline__ uint64_t expand_bits_21(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__host__ __device__ __forceinline__ uint32_t compact_bits_21(uint64_t v)
{
    v &= 0x1249249249249249ull;
    v = (v | (v >> 2)) & 0x10c30c30c30c30c3ull;
    v = (v | (v >> 4)) & 0x100f00f00f00f00full;
    v = (v | (v >> 8)) & 0x001f0000ff0000ffull;
    v = (v | (v >> 16)) & 0x001f00000000ffffull;
    v = (v | (v >> 32)) & 0x1fffffull;
    return (uint32_t)v;
}

template <typename Key> struct Morton;
template <> struct Morton<uint32_t> {
    static constexpr int kMaxBits = 10;
    static constexpr uint32_t kEmpty = 0xffffffffu;
    __host__ __device__ static __forceinline__ uint32_t encode(uint32_t x, uint32_t y, uint32_t z)
    {
        return expand_bits_10(x) | (expand_bits_10(y) << 1) | (expand_bits_10(z) << 2);
    }
    __host__ __device__ static __forceinline__ void decode(uint32_t k, int& x, int& y, int& z)
    {
        x = (int)compact_bits_10(k); y = (int)compact_bits_10(k >> 1); z = (int)compact_bits_10(k >> 2);
    }
    __host__ __device__ static __forceinline__ uint32_t hash(uint32_t k) { return k * 0x9E3779B1u; }
};
template <> struct Morton<uint64_t> {
    static constexpr int kMaxBits = 21;
    static constexpr uint64_t kEmpty = 0xffffffffffffffffull;
    __host__ __device__ static __forceinline__ uint64_t encode(uint32_t x, uint32_t y, uint32_t z)
    {
        return expand_bits_21(x) | (expand_bits_21(y) << 1) | (expand_bits_21(z) << 2);
    }
    __host__ __device__ static __forceinline__ void decode(uint64_t k, int& x, int& y, int& z)
    {
        x = (int)compact_bits_21(k); y = (int)compact_bits_21(k >> 1); z = (int)compact_bits_21(k >> 2);
    }
    __host__ __device__ static __forceinline__ uint32_t hash(uint64_t k)
    {
        return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 32);
    }
};

// ---- neighbour cell in Morton space: +-1 per axis by dilated-integer arithmetic, no decode / encode round trip.
// `valid` is cleared when the neighbour would leave [0, 2^bits) on some axis.
template <typename Key> struct MortonMask;
template <> struct MortonMask<uint32_t> { static constexpr uint32_t kX = 0x09249249u; };
template <> struct MortonMask<uint64_t> { static constexpr uint64_t kX = 0x1249249249249249ull; };

template <typename Key>
__device__ __forceinline__ Key morton_neighbor(Key key, int ox, int oy, int oz, Key key_mask, bool& valid)
{
    Key out = 0;
    const int o[3] = { ox, oy, oz };
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const Key md = (MortonMask<Key>::kX << d) & key_mask;     // bits of axis d that are in use
        const Key part = key & md;
        Key r = part;
        if (o[d] > 0) { valid = valid && (part != md); r = ((key | ~md) + (Key)1) & md; }
        if (o[d] < 0) { valid = valid && (part != 0); r = (part - (Key)1) & md; }
        out |= r;
    }
    return out;
}

// ---- open addressing hash  cell key -> [start, end) of the cell's run in the sorted point array.  One slot is ONE 16-byte
// vector load {key, start, end}, so a neighbour lookup is a single memory round trip (no key -> id -> cell_start chain).
// Empty = all ones.
template <typename Key> struct HashSlot;
template <> struct HashSlot<uint32_t> {
    typedef uint4 Raw;
    __device__ static __forceinline__ bool try_insert(Raw* table, uint32_t slot, uint32_t key, uint32_t start, uint32_t end)
    {
        if (atomicCAS(&table[slot].x, 0xffffffffu, key) != 0xffffffffu) return false;
        table[slot].y = start;
        table[slot].z = end;
        return true;
    }
    __device__ static __forceinline__ Raw load(const Raw* table, uint32_t slot) { return __ldg(table + slot); }
    __device__ static __forceinline__ bool matches(const Raw& e, uint32_t key) { return e.x == key; }
    __device__ static __forceinline__ bool is_empty(const Raw& e) { return e.x == 0xffffffffu; }
    __device__ static __forceinline__ int start(const Raw& e) { return (int)e.y; }
    __device__ static __forceinline__ int count(const Raw& e) { return (int)(e.z - e.y); }
};
template <> struct HashSlot<uint64_t> {
    typedef ulonglong2 Raw;
    __device__ static __forceinline__ bool try_insert(Raw* table, uint32_t slot, uint64_t key, uint32_t start, uint32_t end)
    {
        if (atomicCAS(&table[slot].x, ~0ull, (unsigned long long)key) != ~0ull) return false;
        table[slot].y = (unsigned long long)start | ((unsigned long long)end << 32);
        return true;
    }
    __device__ static __forceinline__ Raw load(const Raw* table, uint32_t slot) { return __ldg(table + slot); }
    __device__ static __forceinline__ bool matches(const Raw& e, uint64_t key) { return e.x == key; }
    __device__ static __forceinline__ bool is_empty(const Raw& e) { return e.x == ~0ull; }
    __device__ static __forceinline__ int start(const Raw& e) { return (int)(uint32_t)e.y; }
    __device__ static __forceinline__ int count(const Raw& e) { return (int)((uint32_t)(e.y >> 32) - (uint32_t)e.y); }
};

// ---- order preserving float <-> uint mapping for atomic min / max
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o)
{
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming (read-once) loads / write-once stores: keep them out of L1 so the query's candidate tiles stay cached
__device__ __forceinline__ float4 ld_nc_f4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

}  // namespace tnsb
