// query_brick.cuh -- the fixed-radius distance query on a HALF-RADIUS grid, a lane owns a QUERY.  Default query path;
// replaces _solve_leaves / _prepare_brute_force[_simd] / _brute_force[_simd] of the reference
// (TreeNSearch.cpp:1823-1872, :2161-2399, :2400-2569).
//
// Grid: cell edge = r_max * (1 + 2^-13) / 2, linear row keys  key = (z * ny + y) * nx + x  (x fastest), prefix table
// first[key] = number of points with a smaller key, sorted points as PLANES x[], y[], z[], id[] (grid_build.cuh).  A neighbour
// of a point of cell (cx, cy, cz) lies in cells [cx-2, cx+2] x [cy-2, cy+2] x [cz-2, cz+2]: 25 rows, and inside a row the cells
// are CONSECUTIVE keys, i.e. one contiguous run of every plane.  15.6 r^3 of candidate volume instead of the 27 r^3 of a
// cell = r grid, and per query the rows -- and the cells inside every row -- that cannot hold a point within r are culled:
// ~75 distance tests per query at ~30 neighbours (2.5 tests per hit; the 27-cell stencil needs 6.5).
//
// Work decomposition (one launch per active ordered pair set_i -> set_j):
//   * brick_plan_kernel cuts the grid into bricks of <= 32 x 4 x 4 cells whose candidate SLAB (the brick plus 2 cells on every
//     side: <= 64 rows of <= 36 cells) fits a shared memory slab buffer; bricks that do not fit are split.
//   * brick_query_kernel: one persistent CTA per SM = NCONS consumer warps + 1 producer warp, two slab buffers.
//     PRODUCER: pulls the next brick from a ticket counter, reads the row boundaries from the prefix table, stages every slab
//       row with one cp.async.bulk per plane (global -> shared through the TMA unit, completion counted in bytes on the buffer's
//       `full` mbarrier) and builds the brick's boundary table T (slab position of every cell boundary) -- while the consumers
//       still work on the previous brick in the other buffer.
//     CONSUMERS: a warp takes 32 consecutive queries of the brick; every LANE owns one query: it culls its 25 rows, turns them
//       into <= 26 (first, last) slab ranges (its own record is cut out of its own row), and walks ALL its ranges in one
//       flattened loop: three conflict-light LDS.32 per candidate (adjacent lanes read adjacent words of a plane), the
//       reference's exact arithmetic d2 = fma(dz,dz, fma(dx,dx, dy*dy)) <= r^2 (TreeNSearch.cpp:2477-2486 as compiled,
//       SURVEY.md §0.5), a hit is ONE predicated 16-bit store into the lane's private column.  No ballot, no popc, no shuffle
//       in the inner loop; lanes that run out of candidates early walk dummy records.  The warp then reserves room for its 32
//       lists with one atomicAdd and writes them as [n, j0, j1, ...] (TreeNSearch.h:395), one coalesced store per 32 words.
//       Warps move on to the next brick on their own (`empty` mbarrier per buffer): no CTA-wide barrier anywhere.
// Queries with more than kMaxTot candidates, lists longer than the column, and cells too dense for any slab take a
// warp-cooperative two-pass slow path that reads the candidates from global memory.
#pragma once
#include "common.cuh"
#include "query.cuh"

namespace tnsb {

constexpr int kBX = 32, kBY = 4, kBZ = 4;                // largest brick, in cells
constexpr int kRowPitch = kBY + 4;                       // slab rows are indexed (z - z0 + 2) * 8 + (y - y0 + 2)
constexpr int kSlabRows = (kBY + 4) * (kBZ + 4);         // 64 candidate rows per brick
constexpr int kQRows = kBY * kBZ;                        // 16 query rows per brick, indexed (z - z0) * 4 + (y - y0)
constexpr int kTW = kBX + 5;                             // cell boundaries per slab row (odd: spreads the rows over the banks)
constexpr int kDummy = 256;                              // dummy records (x = 3e38) behind the x plane
constexpr int kTabH = 32;                                // per-lane range table: 26 ranges + 4 dummy ranges + landing entry + prefetch
constexpr int kMaxTot = 1024;                            // candidates per query on the fast path (variable radii: cells of r_max / 2 hold many candidates of a small query)
constexpr int kColStride = 68;                           // bytes between consecutive hits of one lane (34 uint16: conflict-free column reads)
constexpr uint32_t kBrickSlow = 1u << 24;                // task flag: single cell whose slab does not fit
constexpr uint32_t kSlowPart = 64;                       // queries per task of such a cell: a dense cell becomes many tasks, i.e. many CTAs
constexpr uint32_t kBrickPart = 1u << 25;                // task flag: a HEAVY staged brick (dense cells: its queries all take the warp-cooperative path) is
                                                         // handed out several times, kSlowPart queries per task (part index in z0 >> 21): every CTA stages the
                                                         // (small) slab again, the queries spread over the GPU instead of queueing in one CTA

struct BrickTask {
    int x0, y0, z0;
    uint32_t dims;      // ex | ey << 8 | ez << 16 | flags
};

// sorted points of one set: records (x, y, z, bits(id)), r^2 (variable radius only), prefix cell table
struct BrickSet {
    const float4* pts;
    const float* r2;
    const uint32_t* first;
};

struct BrickArgs {
    BrickGrid g;
    BrickSet q;               // searching set (set_i)
    BrickSet c;               // searched set (set_j)
    int same_set;
    int query_limit;
    float r2_fixed;
    float cull_r2;            // (largest search distance / cell)^2 with 0.2 % slack: rows / cells farther than that are skipped
    float inv_cell_f;
    // plan
    BrickTask* tasks;
    uint32_t max_tasks;
    uint32_t* n_tasks;
    int* plan_overflow;
    uint32_t* ticket;
    // output
    int32_t* ragged;
    long long capacity;
    long long* list_pos;
    unsigned long long* cursor;
    unsigned long long* n_neighbors;
    unsigned long long* n_slow;
    int* max_list;            // longest list written (atomicMax): picks the hit column height of the next run
    int host_out;             // the ragged buffer is mapped host memory: lists leave the SM as aligned, fully coalesced 128-byte stores
    int sort_lists;           // ascending neighbour ids inside every list (the reference's order)
    // lists longer than a warp's scratch that must be sorted AND live in mapped host memory are built and sorted in this device buffer
    // first (a bump allocator per run), then copied out: sorting them in place would be a PCIe round trip per compare-exchange
    int32_t* long_scratch;
    unsigned long long* long_cursor;
    long long long_cap;
    int* overflow;
};

// ---- shared memory geometry (bytes) ---------------------------------------------------------------------------------------
// Per slab buffer: SLAB records + kDummy dummy records (x = 3e38: d2 = inf, never a hit) (+ r^2 of both), boundary table T, meta.
template <int SLAB, int KMAX, bool SYM>
struct BrickSmem {
    static constexpr int kOffSlab = 0;
    static constexpr int kOffR2 = (SLAB + kDummy) * 16;
    static constexpr int kOffT = kOffR2 + (SYM ? (SLAB + kDummy) * 4 : 0);
    static constexpr int kOffMeta = kOffT + ((kSlabRows * kTW * 2 + 15) & ~15);
    // meta words: [0] x0 [1] y0 [2] z0 [3] dims | flags [4] n queries [5] staged [6] next warp task [7] end
    //             [8, 24) qs   [24, 41) qoff   [41] queries per warp task   [42] first query of the task   [44, 60) qdelta
    static constexpr int kMetaWords = 64;
    static constexpr int kOffRowKey = kOffMeta + kMetaWords * 4;        // producer scratch: key of the first cell of every slab row
    static constexpr int kOffRowBase = kOffRowKey + kSlabRows * 4;      //                   slab position - global position of the row's records
    static constexpr int kBufBytes = kOffRowBase + kSlabRows * 4;
    static constexpr int kOffBars = 2 * kBufBytes;                      // full[0], full[1], empty[0], empty[1]
    static constexpr int kOffWarp = kOffBars + 64;
    static constexpr int kTabBytes = kTabH * 128;
    static constexpr int kWarpBytes = (kTabBytes + (KMAX + 1) * kColStride + 15) & ~15;
    static constexpr int total(int n_cons) { return kOffWarp + n_cons * kWarpBytes; }
    static_assert((SLAB + kDummy) * 16 < 65536, "slab offsets are stored as 16-bit byte offsets");
};

// ---- shared memory / mbarrier / bulk copy primitives (32-bit shared window addresses) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (uint32_t)v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BRICK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"      // suspend-time hint: the thread sleeps until the phase completes instead of spinning through issue slots
        "@p bra BRICK_DONE;\n"
        "bra BRICK_WAIT;\n"
        "BRICK_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (UBLKCP); completion is signalled on the mbarrier as transferred bytes.
// dst / src 16-byte aligned, bytes a non-zero multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// -1 if a >= b else 0 (FSET: no predicate round trip)
__device__ __forceinline__ int fge_mask(float a, float b)
{
    int d;
    asm("set.ge.s32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b));
    return d;
}
// streaming 32-bit store to base[off] / base[off + 32] (one mad.wide for the address)
__device__ __forceinline__ void stg_cs_at(const int32_t* base, uint32_t off, int v)
{
    asm volatile("{\n.reg .u64 a;\nmad.wide.u32 a, %1, 4, %0;\nst.global.cs.s32 [a], %2;\n}\n" ::"l"(base), "r"(off), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_cs_at32(const int32_t* base, uint32_t off, int v)
{
    asm volatile("{\n.reg .u64 a;\nmad.wide.u32 a, %1, 4, %0;\nst.global.cs.s32 [a+128], %2;\n}\n" ::"l"(base), "r"(off), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// cell coordinate exactly as the key generation computes it (grid_build.cuh brick_keygen_count_kernel): fp64, floor, clamp
__device__ __forceinline__ int brick_cell(float v, double bottom, double inv_cell, int n, double& t)
{
    t = ((double)v - bottom) * inv_cell;
    return min(max(__double2int_rd(t), 0), n - 1);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Plan: one warp per 32 x 4 x 4 brick; bricks without queries are dropped, bricks whose candidate slab exceeds slab_cap are
// split (x first: rows stay long) down to single cells, which are flagged for the slow path.
__device__ __forceinline__ uint32_t plan_rows(const uint32_t* first, const BrickGrid& g, int xa, int xb, int y0, int ny_rows, int z0, int nz_rows, int lane,
                                              uint32_t* row_max = nullptr)
{
    uint32_t s = 0, m = 0;
    const int n_rows = ny_rows * nz_rows;
    for (int r = lane; r < n_rows; r += 32) {
        const int y = y0 + r % ny_rows, z = z0 + r / ny_rows;
        if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) continue;
        const uint32_t key0 = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
        const uint32_t c = first[key0 + xb] - first[key0 + xa];
        s += c;
        m = max(m, c);
    }
    if (row_max) *row_max = __reduce_max_sync(kFull, m);
    return __reduce_add_sync(kFull, s);
}

__global__ void __launch_bounds__(256) brick_plan_kernel(const BrickGrid g, const uint32_t* __restrict__ q_first, const uint32_t* __restrict__ c_first, int slab_cap,
                                                         BrickTask* __restrict__ tasks, uint32_t max_tasks, uint32_t* __restrict__ n_tasks, int* __restrict__ plan_overflow)
{
    __shared__ int s_stack[8][16][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbx = ceil_div(g.nx, kBX), nby = ceil_div(g.ny, kBY), nbz = ceil_div(g.nz, kBZ);
    const long long n_bricks = (long long)nbx * nby * nbz;
    int(*stack)[6] = s_stack[warp];
    for (long long b = (long long)blockIdx.x * 8 + warp; b < n_bricks; b += (long long)gridDim.x * 8) {
        const int bx = (int)(b % nbx), by = (int)((b / nbx) % nby), bz = (int)(b / ((long long)nbx * nby));
        if (lane == 0) {
            stack[0][0] = bx * kBX; stack[0][1] = by * kBY; stack[0][2] = bz * kBZ;
            stack[0][3] = min(kBX, g.nx - bx * kBX); stack[0][4] = min(kBY, g.ny - by * kBY); stack[0][5] = min(kBZ, g.nz - bz * kBZ);
        }
        int sp = 1;
        __syncwarp();
        while (sp > 0) {
            sp--;
            const int x0 = stack[sp][0], y0 = stack[sp][1], z0 = stack[sp][2], ex = stack[sp][3], ey = stack[sp][4], ez = stack[sp][5];
            __syncwarp();
            const uint32_t nq = plan_rows(q_first, g, x0, x0 + ex, y0, ey, z0, ez, lane);
            if (nq == 0) continue;
            uint32_t row_max = 0;
            const uint32_t nc = plan_rows(c_first, g, max(x0 - 2, 0), min(x0 + ex + 2, g.nx), y0 - 2, ey + 4, z0 - 2, ez + 4, lane, &row_max);
            uint32_t flags = 0;
            if (nc > (uint32_t)slab_cap) {
                if (ex > 1 || ey > 1 || ez > 1) {
                    if (lane == 0) {
                        int* a = stack[sp];
                        int* c = stack[sp + 1];
                        for (int k = 0; k < 6; k++) c[k] = a[k];
                        if (ex > 1) { const int h = ex >> 1; a[3] = h; c[0] = x0 + h; c[3] = ex - h; }
                        else if (ey > 1) { const int h = ey >> 1; a[4] = h; c[1] = y0 + h; c[4] = ey - h; }
                        else { const int h = ez >> 1; a[5] = h; c[2] = z0 + h; c[5] = ez - h; }
                    }
                    sp += 2;
                    __syncwarp();
                    continue;
                }
                flags = kBrickSlow;
            }
            if (lane == 0) {
                // an unstageable cell is cut into tasks of kSlowPart queries (its queries read their candidates from global memory, nothing
                // is shared through the slab): dims = 1 | part << 8 | kBrickSlow
                // a staged brick is heavy when its queries will mostly exceed kMaxTot candidates: cells so full that an average query does
                // (125 cells per query), or one row of the slab that alone holds hundreds of records (a blob inside an otherwise sparse brick)
                const uint32_t slab_cells = (uint32_t)((ex + 4) * (ey + 4) * (ez + 4));
                const bool heavy = !flags && nq > kSlowPart && ((unsigned long long)nc * 125ull > (unsigned long long)kMaxTot * slab_cells || row_max >= 400u);
                const uint32_t parts = flags ? min((nq + kSlowPart - 1u) / kSlowPart, 0xffffu) : (heavy ? min((nq + kSlowPart - 1u) / kSlowPart, 0x7ffu) : 1u);
                const uint32_t id = atomicAdd(n_tasks, parts);
                for (uint32_t p = 0; p < parts; p++) {
                    if (id + p < max_tasks) {
                        BrickTask t;
                        t.x0 = x0; t.y0 = y0; t.z0 = heavy ? (int)((uint32_t)z0 | (p << 21)) : z0;
                        t.dims = flags ? (1u | (p << 8) | flags) : ((uint32_t)ex | ((uint32_t)ey << 8) | ((uint32_t)ez << 16) | (heavy ? kBrickPart : 0u));
                        tasks[id + p] = t;
                    } else {
                        *plan_overflow = 1;
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ascending sort of n ints by one warp: normalized bitonic network (every compare-exchange puts the smaller value at the lower
// index, so the virtual +inf padding up to the next power of two never moves).  Used by the slow paths only.
template <typename Load, typename Store>
__device__ __forceinline__ void warp_bitonic_sort(int n, int lane, Load ld, Store st)
{
    if (n < 2) return;
    int lg = 1;
    while ((1 << lg) < n) lg++;
    const int half = 1 << (lg - 1);                              // pairs per step
    for (int s = 1; s <= lg; s++) {
        for (int j = s - 1; j >= 0; j--) {
            for (int t = lane; t < half; t += 32) {
                const int lo = ((t >> j) << (j + 1)) | (t & ((1 << j) - 1));
                const int hi = (j == s - 1) ? (lo ^ ((1 << s) - 1)) : (lo + (1 << j));
                if (hi < n) {
                    const int x = ld(lo), y = ld(hi);
                    if (x > y) { st(lo, y); st(hi, x); }
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// where the second pass of a slow path writes a list of n ids that did not fit the warp's scratch: straight behind its count word in the
// ragged buffer, or -- sorted lists in mapped host memory -- into a piece of the device scratch (nullptr: none left, sort in place)
__device__ __forceinline__ int32_t* long_list_target(const BrickArgs& a, int32_t* dst, int n, int lane, bool& staged_out)
{
    staged_out = false;
    if (!(a.host_out && a.sort_lists && a.long_scratch)) return dst + 1;
    unsigned long long at = 0;
    if (lane == 0) at = atomicAdd(a.long_cursor, (unsigned long long)n);
    at = __shfl_sync(kFull, at, 0);
    if ((long long)(at + (unsigned long long)n) > a.long_cap) return dst + 1;
    staged_out = true;
    return a.long_scratch + at;
}

// sorts the n ids at `wr` (device memory) and, when they were staged, copies them behind the count word at dst
__device__ __forceinline__ void long_list_finish(const BrickArgs& a, int32_t* dst, int32_t* wr, int n, int lane, bool staged_out)
{
    __syncwarp();
    if (a.sort_lists && n > 1) {
        volatile int32_t* v = wr;
        warp_bitonic_sort(n, lane, [&](int i) { return (int)v[i]; }, [&](int i, int x) { v[i] = x; });
    }
    if (staged_out) {
        __syncwarp();
        for (int i = lane; i < n; i += 32) dst[1 + i] = wr[i];
    }
}

// slow path: one query, the whole warp, candidates from global memory.  The 25 row bounds are fetched by 25 lanes at once, the rows are
// walked with FOUR warp-wide candidate loads in flight (the path is bound by memory latency, not by arithmetic), and the hits of the
// first pass are kept in the warp's scratch: a list that fits it (<= scratch_cap ids) is written out -- sorted if asked -- without a second
// pass; longer lists take a second pass straight into the ragged buffer.
template <bool SYMMETRIC>
__device__ __noinline__ void brick_slow_query(const BrickArgs& a, float qx, float qy, float qz, int qid, float r2, int cx, int cy, int cz, int lane, unsigned& nb_sum,
                                              uint32_t scratch_a, int scratch_cap)
{
    const unsigned lt = lanemask_lt();
    uint32_t lo = 0, hi = 0;
    if (lane < 25) {
        const int y = cy + lane % 5 - 2, z = cz + lane / 5 - 2;
        if (y >= 0 && y < a.g.ny && z >= 0 && z < a.g.nz) {
            const uint32_t key0 = ((uint32_t)z * (uint32_t)a.g.ny + (uint32_t)y) * (uint32_t)a.g.nx;
            lo = a.c.first[key0 + max(cx - 2, 0)];
            hi = a.c.first[key0 + min(cx + 3, a.g.nx)];
        }
    }
    int32_t* dst = nullptr;
    int32_t* wr = nullptr;
    bool staged_out = false;
    int n_list = 0;
    for (int pass = 0; pass < 2; pass++) {
        int n = 0;
        for (int row = 0; row < 25; row++) {
            const uint32_t lo_r = __shfl_sync(kFull, lo, row), hi_r = __shfl_sync(kFull, hi, row);
            for (uint32_t t0 = lo_r; t0 < hi_r; t0 += 128u) {
                float4 v[4];
                float w[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t t = t0 + 32u * u + (uint32_t)lane;
                    v[u] = make_float4(3.0e38f, 0.0f, 0.0f, __int_as_float(-1));
                    w[u] = -1.0f;
                    if (t < hi_r) {
                        v[u] = a.c.pts[t];
                        if (SYMMETRIC) w[u] = a.c.r2[t];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (t0 + 32u * u >= hi_r) break;
                    const int id = __float_as_int(v[u].w);
                    const float d2 = dist2(qx, qy, qz, v[u].x, v[u].y, v[u].z);
                    bool hit = d2 <= r2;
                    if (SYMMETRIC) hit = hit || (d2 <= w[u]);
                    if (a.same_set && id == qid) hit = false;
                    const unsigned m = __ballot_sync(kFull, hit);
                    if (hit) {
                        const int k = n + __popc(m & lt);
                        if (pass == 1) wr[k] = id;
                        else if (k < scratch_cap) sts_u32(scratch_a + (uint32_t)k * 4u, (uint32_t)id);
                    }
                    n += __popc(m);
                }
            }
        }
        if (pass == 0) {
            const unsigned long long need = (unsigned long long)((n + 1 + 3) & ~3);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.cursor, need);
            base = __shfl_sync(kFull, base, 0);
            if ((long long)(base + need) > a.capacity) {
                if (lane == 0) *a.overflow = 1;
                return;
            }
            dst = a.ragged + base;
            if (lane == 0) {
                dst[0] = n;
                a.list_pos[qid] = (long long)base;
            }
            nb_sum += (unsigned)n;
            n_list = n;
            if (n <= scratch_cap) {
                // the whole list sits in the scratch: sort it there, one coalesced copy, done
                __syncwarp();
                if (a.sort_lists && n > 1)
                    warp_bitonic_sort(n, lane, [&](int i) { return (int)lds_u32(scratch_a + (uint32_t)i * 4u); }, [&](int i, int x) { sts_u32(scratch_a + (uint32_t)i * 4u, (uint32_t)x); });
                for (int i = lane; i < n; i += 32) dst[1 + i] = (int)lds_u32(scratch_a + (uint32_t)i * 4u);
                __syncwarp();
                return;
            }
            wr = long_list_target(a, dst, n, lane, staged_out);
        }
    }
    // longer than the warp's scratch (thousands of neighbours): sorted in device memory
    long_list_finish(a, dst, wr, n_list, lane, staged_out);
}


// slow path of a STAGED brick (the list does not fit the hit column, or the query has more than kMaxTot candidates): one query, the
// whole warp, candidates from the slab in shared memory (all 25 rows, +-2 cells, no culling).  The hits of the first pass are kept in
// the warp's scratch: a list that fits it is written out -- sorted if asked -- without a second pass.
template <bool SYMMETRIC>
__device__ __noinline__ void brick_slow_query_staged(const BrickArgs& a, uint32_t slab_a, uint32_t r2_a, const uint16_t* tq, float qx, float qy, float qz, int qid,
                                                    float r2, int self_off, int lane, unsigned& nb_sum, uint32_t scratch_a, int scratch_cap)
{
    const unsigned lt = lanemask_lt();
    int32_t* dst = nullptr;
    int32_t* wr = nullptr;
    bool staged_out = false;
    int n_list = 0;
    for (int pass = 0; pass < 2; pass++) {
        int n = 0;
        for (int row = 0; row < 25; row++) {
            const uint16_t* trow = tq + ((row / 5) * kRowPitch + (row % 5)) * kTW;
            const uint32_t lo = trow[0], hi = trow[5];                // byte offsets of cells cx - 2 .. cx + 2 of this row
            for (uint32_t t0 = lo; t0 < hi; t0 += 32u * 16u) {
                const uint32_t t = t0 + (uint32_t)lane * 16u;
                bool hit = false;
                int id = -1;
                if (t < hi && (int)t != self_off) {
                    const float4 v = lds_f4(slab_a + t);
                    id = __float_as_int(v.w);
                    const float d2 = dist2(qx, qy, qz, v.x, v.y, v.z);
                    hit = d2 <= r2;
                    if (SYMMETRIC) hit = hit || (d2 <= lds_f32(r2_a + (t >> 2)));
                }
                const unsigned m = __ballot_sync(kFull, hit);
                if (hit) {
                    const int k = n + __popc(m & lt);
                    if (pass == 1) wr[k] = id;
                    else if (k < scratch_cap) sts_u32(scratch_a + (uint32_t)k * 4u, (uint32_t)id);
                }
                n += __popc(m);
            }
        }
        if (pass == 0) {
            const unsigned long long need = (unsigned long long)((n + 1 + 3) & ~3);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.cursor, need);
            base = __shfl_sync(kFull, base, 0);
            if ((long long)(base + need) > a.capacity) {
                if (lane == 0) *a.overflow = 1;
                return;
            }
            dst = a.ragged + base;
            if (lane == 0) {
                dst[0] = n;
                a.list_pos[qid] = (long long)base;
            }
            nb_sum += (unsigned)n;
            n_list = n;
            if (n <= scratch_cap) {
                // the whole list sits in the scratch: sort it there, one coalesced copy, done
                __syncwarp();
                if (a.sort_lists && n > 1)
                    warp_bitonic_sort(n, lane, [&](int i) { return (int)lds_u32(scratch_a + (uint32_t)i * 4u); }, [&](int i, int x) { sts_u32(scratch_a + (uint32_t)i * 4u, (uint32_t)x); });
                for (int i = lane; i < n; i += 32) dst[1 + i] = (int)lds_u32(scratch_a + (uint32_t)i * 4u);
                __syncwarp();
                return;
            }
            wr = long_list_target(a, dst, n, lane, staged_out);
        }
    }
    // longer than the warp's scratch (thousands of neighbours): sorted in device memory
    long_list_finish(a, dst, wr, n_list, lane, staged_out);
}

// ---------------------------------------------------------------------------------------------------------------------------
// warps 0 .. NCONS-1 consume, warp NCONS + b produces into slab buffer b
template <int NCONS, int SLAB, int KMAX, bool VARIABLE, bool SYMMETRIC>
__global__ void __launch_bounds__((NCONS + 2) * 32, 1) brick_query_kernel(const BrickArgs a)
{
    typedef BrickSmem<SLAB, KMAX, SYMMETRIC> SM;
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t s_base = smem_u32(s_raw);
    asm volatile("" : "+r"(s_base));                      // opaque: keeps the shared window base in a register instead of re-deriving it per access
    const uint32_t bars = s_base + SM::kOffBars;          // full[b] = bars + 8 b, empty[b] = bars + 16 + 8 b
    const BrickGrid g = a.g;

    // dummy records of both buffers, barriers
    for (int k = tid; k < 2 * kDummy; k += (NCONS + 2) * 32) {
        const int b = k / kDummy, i = k % kDummy;
        reinterpret_cast<float4*>(s_raw + b * SM::kBufBytes + SM::kOffSlab)[SLAB + i] = make_float4(3.0e38f, 0.0f, 0.0f, __int_as_float(-1));
        if (SYMMETRIC) reinterpret_cast<float*>(s_raw + b * SM::kBufBytes + SM::kOffR2)[SLAB + i] = -1.0f;
    }
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 8, 1);
        mbar_init(bars + 16, NCONS);
        mbar_init(bars + 24, NCONS);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t n_tasks = min(*a.n_tasks, a.max_tasks);

    if (warp >= NCONS) {
        // =============================================== PRODUCER of buffer b ===============================================
        const uint32_t b = (uint32_t)(warp - NCONS);
        unsigned char* const buf = s_raw + b * SM::kBufBytes;
        const uint32_t buf_a = s_base + b * SM::kBufBytes;
        uint32_t* const meta = reinterpret_cast<uint32_t*>(buf + SM::kOffMeta);
        uint32_t* const s_rowkey = reinterpret_cast<uint32_t*>(buf + SM::kOffRowKey);
        int* const s_rowbase = reinterpret_cast<int*>(buf + SM::kOffRowBase);
        uint16_t* const sT = reinterpret_cast<uint16_t*>(buf + SM::kOffT);
        const uint32_t full = bars + 8 * b, empty = bars + 16 + 8 * b;
        for (uint32_t use = 0;; use++) {
            // ---- everything that only needs registers happens BEFORE the wait for the buffer: next ticket, brick, row boundaries
            uint32_t task = 0;
            if (lane == 0) task = atomicAdd(a.ticket, 1u);
            task = __shfl_sync(kFull, task, 0);
            if (task >= n_tasks) {
                if (use >= 1) mbar_wait(empty, (use - 1u) & 1u);
                if (lane == 0) {
                    meta[7] = 1u;
                    mbar_arrive(full);
                }
                break;
            }
            const BrickTask bt = a.tasks[task];
            const bool part_brick = (bt.dims & kBrickPart) != 0;
            const int x0 = bt.x0, y0 = bt.y0, z0 = part_brick ? (int)((uint32_t)bt.z0 & 0x1fffffu) : bt.z0;
            const uint32_t brick_part = part_brick ? ((uint32_t)bt.z0 >> 21) : 0u;
            const bool slow_brick = (bt.dims & kBrickSlow) != 0;
            const int ex = (int)(bt.dims & 0xffu), ey = slow_brick ? 1 : (int)((bt.dims >> 8) & 0xffu), ez = slow_brick ? 1 : (int)((bt.dims >> 16) & 0xffu);
            const uint32_t slow_part = (bt.dims >> 8) & 0xffffu;          // unstageable cell: which kSlowPart queries of it
            const int tw = ex + 5;
            const int xa = max(x0 - 2, 0), xb = min(x0 + ex + 2, g.nx);
            // slab rows: rows lane and lane + 32 (row = sz * 8 + sy)
            uint32_t len[2], g0[2], rkey[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = lane + 32 * h;
                const int sy = r & 7, sz = r >> 3;
                const int y = y0 - 2 + sy, z = z0 - 2 + sz;
                len[h] = 0; g0[h] = 0;
                rkey[h] = 0xffffffffu;
                if (!slow_brick && sy < ey + 4 && sz < ez + 4 && y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
                    rkey[h] = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
                    g0[h] = a.c.first[rkey[h] + xa];
                    len[h] = a.c.first[rkey[h] + xb] - g0[h];
                    if (len[h] == 0) rkey[h] = 0xffffffffu;
                }
            }
            // query rows (row = rz * 4 + ry)
            uint32_t qs = 0, qcnt = 0;
            {
                const int ry = lane & 3, rz = lane >> 2;
                if (lane < kQRows && ry < ey && rz < ez) {
                    const uint32_t key0 = ((uint32_t)(z0 + rz) * (uint32_t)g.ny + (uint32_t)(y0 + ry)) * (uint32_t)g.nx;
                    qs = a.q.first[key0 + x0];
                    qcnt = a.q.first[key0 + x0 + ex] - qs;
                    if (slow_brick) {
                        const uint32_t o = min(qcnt, slow_part * kSlowPart);
                        qs += o;
                        qcnt = slow_part == 0xfffeu ? qcnt - o : min(qcnt - o, kSlowPart);     // the last representable part (the planner emits at most 0xffff) takes the rest
                    }
                }
            }
            const uint32_t inc0 = (uint32_t)warp_inclusive_scan((int)len[0], lane);
            const uint32_t tot0 = __shfl_sync(kFull, inc0, 31);
            const uint32_t inc1 = (uint32_t)warp_inclusive_scan((int)len[1], lane) + tot0;
            const uint32_t total = __shfl_sync(kFull, inc1, 31);
            const uint32_t ro[2] = { inc0 - len[0], inc1 - len[1] };
            const bool staged = !slow_brick && total <= (uint32_t)SLAB;
            const uint32_t qinc = (uint32_t)warp_inclusive_scan((int)qcnt, lane);
            const uint32_t nq_all = __shfl_sync(kFull, qinc, 31);

            if (use >= 1) mbar_wait(empty, (use - 1u) & 1u);          // every consumer warp has left the brick that used this buffer
            s_rowkey[lane] = rkey[0];
            s_rowkey[lane + 32] = rkey[1];
            s_rowbase[lane] = (int)ro[0] - (int)g0[0];
            s_rowbase[lane + 32] = (int)ro[1] - (int)g0[1];
            if (lane < kQRows) { meta[8 + lane] = qs; meta[24 + lane] = qinc - qcnt; }
            if (lane == kQRows) meta[24 + lane] = nq_all;
            if (lane == 0) {
                meta[0] = (uint32_t)x0; meta[1] = (uint32_t)y0; meta[2] = (uint32_t)z0; meta[3] = bt.dims;
                // queries [q_lo, q_hi) of the brick belong to this task (all of them, or one part of a heavy brick)
                const uint32_t q_lo = part_brick ? min(nq_all, brick_part * kSlowPart) : 0u;
                const uint32_t q_hi = part_brick ? (brick_part == 0x7feu ? nq_all : min(nq_all, q_lo + kSlowPart)) : nq_all;
                const uint32_t nq_task = q_hi - q_lo;
                meta[4] = q_hi; meta[42] = q_lo; meta[5] = staged ? 1u : 0u; meta[6] = 0u; meta[7] = 0u;
                // queries per warp task: 32 -- or, for a brick with few queries and a heavy slab (dense neighbourhoods: long walks, lists
                // that overflow into the warp-cooperative paths), as few as it takes to give every consumer warp a share of the brick:
                // only two bricks are in flight per SM, so a brick that is one task would leave all other consumer warps idle
                uint32_t tq = 32u;
                if ((slow_brick || (total >= 1024u && total >= 8u * nq_task)) && nq_task < 32u * (uint32_t)NCONS) tq = max((nq_task + (uint32_t)NCONS - 1u) / (uint32_t)NCONS, 1u);
                meta[41] = tq;
            }
            __syncwarp();
            if (lane < kQRows) meta[44 + lane] = (uint32_t)s_rowbase[((lane >> 2) + 2) * kRowPitch + (lane & 3) + 2];
            if (staged) {
                // ---- the copies first (they run while the boundary table is built), then T; the producer's ARRIVAL on `full` comes
                // last (release: meta and T are visible to a consumer that sees the phase complete)
                if (lane == 0) {
                    fence_proxy_async();    // the consumers' generic reads of this buffer (ordered by the empty barrier) precede the async writes
                    mbar_expect_tx(full, total * 16u);
                }
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (len[h] == 0) continue;
                    bulk_g2s(buf_a + SM::kOffSlab + ro[h] * 16u, a.c.pts + g0[h], len[h] * 16u, full);
                }
                if (SYMMETRIC) {
                    // r^2 of the rows' records (4-byte elements have no 16-byte alignment to offer a bulk copy): 4-byte asynchronous copies,
                    // all of them in flight at once; the producer waits for them once, right before its arrival on `full`
                    float* const sr2 = reinterpret_cast<float*>(buf + SM::kOffR2);
#pragma unroll 4
                    for (int r = 0; r < kSlabRows; r++) {
                        const uint32_t lr = __shfl_sync(kFull, len[r >> 5], r & 31), gr = __shfl_sync(kFull, g0[r >> 5], r & 31), rr_ = __shfl_sync(kFull, ro[r >> 5], r & 31);
                        for (uint32_t k = lane; k < lr; k += 32) cp_async_4(sr2 + rr_ + k, a.c.r2 + gr + k);
                    }
                    cp_async_commit();
                }
                // T[row][i] = slab position of the first record of cell x0 - 2 + i of that row: one row per load instruction (lane = i),
                // sixteen rows in flight
                const int n_rows = (ez + 4) * kRowPitch;
                const int xi0 = min(max(x0 - 2 + lane, 0), g.nx), xi1 = min(max(x0 - 2 + lane + 32, 0), g.nx);
                for (int r0 = 0; r0 < n_rows; r0 += 16) {
                    uint32_t v0[16], v1[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const uint32_t key0 = s_rowkey[r0 + j];
                        v0[j] = 0; v1[j] = 0;
                        if (key0 != 0xffffffffu) {
                            v0[j] = a.c.first[key0 + (uint32_t)xi0];
                            if (lane + 32 < tw) v1[j] = a.c.first[key0 + (uint32_t)xi1];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int r = r0 + j;
                        const bool ok = s_rowkey[r] != 0xffffffffu;
                        const int rb = s_rowbase[r];
                        sT[r * kTW + lane] = ok ? (uint16_t)(((int)v0[j] + rb) * 16) : (uint16_t)0;
                        if (lane + 32 < kTW) sT[r * kTW + lane + 32] = (ok && lane + 32 < tw) ? (uint16_t)(((int)v1[j] + rb) * 16) : (uint16_t)0;
                    }
                }
            }
            if (SYMMETRIC) cp_async_wait_all();
            __syncwarp();
            if (lane == 0) mbar_arrive(full);
        }
    } else {
        // =============================================== CONSUMERS ===============================================
        const uint32_t tab_a = s_base + SM::kOffWarp + warp * SM::kWarpBytes + lane * 4;      // this lane's column of the range table
        const uint32_t col_w = s_base + SM::kOffWarp + warp * SM::kWarpBytes + SM::kTabBytes;
        const uint32_t col_a = col_w + lane * 2;                                                // this lane's hit column
        const uint32_t col_cap = col_a + (uint32_t)KMAX * kColStride;
        const bool same_set = a.same_set != 0;
        const int query_limit = a.query_limit;
        constexpr uint32_t kDummyOff = (uint32_t)SLAB * 16u;
        unsigned nb_sum = 0, slow_sum = 0;
        int n_max = 0;

        // one brick out of slab buffer B (B is a compile time constant: shared memory addresses fold into the load instructions);
        // returns false when the producer of this buffer has signalled the end of the task list
        auto consume = [&](auto btag, uint32_t use) -> bool {
            constexpr int B = decltype(btag)::value;
            unsigned char* const buf = s_raw + B * SM::kBufBytes;
            const uint32_t buf_a = s_base + B * SM::kBufBytes;
            uint32_t* const meta = reinterpret_cast<uint32_t*>(buf + SM::kOffMeta);
            const uint16_t* const sT = reinterpret_cast<const uint16_t*>(buf + SM::kOffT);
            mbar_wait(bars + 8 * B, use & 1u);                  // tables written, slab rows landed
            if (meta[7]) return false;
            const int x0 = (int)meta[0], y0 = (int)meta[1], z0 = (int)meta[2];
            const int ex = (int)(meta[3] & 0xffu);
            const bool staged = meta[5] != 0u;
            const int nq = (int)meta[4];           // end of this task's queries (brick-local numbering)
            const int q_lo = (int)meta[42];        // their start (0 unless the task is one part of a heavy brick)
            const int tqs = (int)meta[41];
            const uint32_t slab_a = buf_a + SM::kOffSlab;

            for (;;) {
                int wt = 0;
                if (lane == 0) wt = (int)atomicAdd(&meta[6], 1u);
                wt = __shfl_sync(kFull, wt, 0);
                if (q_lo + wt * tqs >= nq) break;
                const int ci = q_lo + wt * tqs + lane;
                const bool has = lane < tqs && ci < nq;
                // query row of this lane: last row whose first query is <= ci (binary search over the 16 row offsets)
                int rr = (ci >= (int)meta[24 + 8]) ? 8 : 0;
                rr += (ci >= (int)meta[24 + rr + 4]) ? 4 : 0;
                rr += (ci >= (int)meta[24 + rr + 2]) ? 2 : 0;
                rr += (ci >= (int)meta[24 + rr + 1]) ? 1 : 0;
                const int ry = rr & 3, rz = rr >> 2;
                const int qp = has ? (int)meta[8 + rr] + (ci - (int)meta[24 + rr]) : 0;
                float4 q = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0x7fffffff));
                float r2 = a.r2_fixed;
                if (has) {
                    if (same_set && staged) {
                        // the query's own row is part of the staged slab: its record is read from shared memory (a global load here is a
                        // full L2 / DRAM round trip at the head of every task, with 16 warps per SM to hide it)
                        const uint32_t qoff = (uint32_t)(qp + (int)meta[44 + rr]) * 16u;
                        q = lds_f4(slab_a + qoff);
                        if (VARIABLE) r2 = SYMMETRIC ? lds_f32(buf_a + SM::kOffR2 + (qoff >> 2)) : a.q.r2[qp];
                    } else {
                        q = a.q.pts[qp];
                        if (VARIABLE) r2 = a.q.r2[qp];
                    }
                }
                const int qid = __float_as_int(q.w);
                const bool active = has && qid < query_limit;
                if (!__any_sync(kFull, active)) continue;          // a task of find-only points (the halo layers of a Z-slab shard)
                double tx, ty, tz;
                const int cx = brick_cell(q.x, g.bottom[0], g.inv_cell, g.nx, tx);
                (void)brick_cell(q.y, g.bottom[1], g.inv_cell, g.ny, ty);
                (void)brick_cell(q.z, g.bottom[2], g.inv_cell, g.nz, tz);
                const int cy = y0 + ry, cz = z0 + rz;
                bool slow = active && !staged;
                uint32_t total = 0;
                uint32_t tp = tab_a;                                // next free entry of this lane's range table
                if (staged) {
                    // ---- this lane's candidate ranges: rows within the search distance, cells culled per row
                    const float fx = (float)(tx - (double)cx), fy = (float)(ty - (double)cy), fz = (float)(tz - (double)cz);
                    float cull = a.cull_r2;
                    if (VARIABLE && !SYMMETRIC) {
                        const float rc = __fmul_rn(sqrtf(r2), a.inv_cell_f);
                        cull = __fmul_rn(__fmul_rn(rc, rc), 1.002f);
                    }
                    if (!active) cull = -1.0f;
                    const int ix = min(max(cx - x0, 0), ex - 1);
                    // squared distance of the query to the slabs of cells at offset -2, -1, 0, +1, +2 along y and z ...
                    float dy2[5], dz2[5];
                    dy2[0] = fy + 1.0f; dy2[1] = fy; dy2[2] = 0.0f; dy2[3] = 1.0f - fy; dy2[4] = 2.0f - fy;
                    dz2[0] = fz + 1.0f; dz2[1] = fz; dz2[2] = 0.0f; dz2[3] = 1.0f - fz; dz2[4] = 2.0f - fz;
#pragma unroll
                    for (int k = 0; k < 5; k++) { dy2[k] = dy2[k] * dy2[k]; dz2[k] = dz2[k] * dz2[k]; }
                    // ... and to the cells at x offset -2, -1, +1, +2: cell k of a row is needed iff dx2[k] + d2yz <= cull
                    const float ax0 = (fx + 1.0f) * (fx + 1.0f), ax1 = fx * fx, ax2 = (1.0f - fx) * (1.0f - fx), ax3 = (2.0f - fx) * (2.0f - fx);
                    const int self_pos = same_set ? (qp + (int)meta[44 + rr]) * 16 : -1;        // T holds byte offsets into the slab
                    const uint16_t* const tq = sT + (rz * kRowPitch + ry) * kTW + ix;
#pragma unroll
                    for (int dz = 0; dz < 5; dz++) {
#pragma unroll
                        for (int dy = 0; dy < 5; dy++) {
                            const float rem = cull - (dy2[dy] + dz2[dz]);
                            const int i_lo = 2 + fge_mask(rem, ax1) + fge_mask(rem, ax0);
                            const int i_hi = 3 - fge_mask(rem, ax2) - fge_mask(rem, ax3);
                            const uint16_t* const trow = tq + (dz * kRowPitch + dy) * kTW;
                            uint32_t lo = trow[i_lo];
                            const uint32_t hi = trow[i_hi];
                            const bool row_ok = rem >= 0.0f;
                            if (dy == 2 && dz == 2 && row_ok && self_pos >= (int)lo && self_pos < (int)hi) {
                                // own row: the query's own record is cut out (only the identical (set, index) is excluded, TreeNSearch.cpp:2464-2466)
                                if (self_pos > (int)lo) {
                                    sts_u32(tp, lo | ((uint32_t)self_pos << 16));
                                    tp += 128u;
                                    total += (uint32_t)self_pos - lo;
                                }
                                lo = (uint32_t)self_pos + 16u;
                            }
                            if (row_ok && hi > lo) {
                                sts_u32(tp, lo | (hi << 16));
                                tp += 128u;
                                total += hi - lo;
                            }
                        }
                    }
                    total >>= 4;        // bytes -> records
                    if (total > (uint32_t)kMaxTot) { slow = true; total = 0; tp = tab_a; }
                }
                const int maxtot = (int)__reduce_max_sync(kFull, total);
                uint32_t ca = col_a;
                if (maxtot > 0) {
                    // lanes with fewer candidates walk the dummy records (never a hit) until the longest lane is done
                    int pad = maxtot - (int)total;
                    while (pad > 0) {
                        const int c = min(pad, kDummy);
                        sts_u32(tp, kDummyOff | ((kDummyOff + (uint32_t)c * 16u) << 16));
                        tp += 128u;
                        pad -= c;
                    }
                    sts_u32(tp, kDummyOff | ((kDummyOff + (uint32_t)kDummy * 16u) << 16));      // landing entry of the last advance

                    // ---- the flattened candidate walk
                    uint32_t w = lds_u32(tab_a);
                    uint32_t p = w & 0xffffu, e = w >> 16;
                    uint32_t wn = lds_u32(tab_a + 128);
                    uint32_t rp = tab_a + 256;
                    // candidates are loaded two iterations ahead of their test
                    auto advance = [&]() {
                        p += 16u;
                        if (p == e) {
                            p = wn & 0xffffu;
                            e = wn >> 16;
                            wn = lds_u32(rp);
                            rp += 128u;
                        }
                    };
                    uint32_t pc0 = p;
                    float4 c0 = lds_f4(slab_a + p);
                    advance();
                    uint32_t pc1 = p;
                    float4 c1 = lds_f4(slab_a + p);
                    advance();
#pragma unroll 6
                    for (int it = 0; it < maxtot; it++) {
                        const uint32_t pn = p;
                        const float4 cn = lds_f4(slab_a + p);
                        advance();
                        const float d2 = dist2(q.x, q.y, q.z, c0.x, c0.y, c0.z);
                        bool hit = d2 <= r2;
                        if (SYMMETRIC) hit = hit || (d2 <= lds_f32(buf_a + SM::kOffR2 + (pc0 >> 2)));
                        if (hit) {
                            sts_u16(ca, pc0);
                            ca = min(ca + (uint32_t)kColStride, col_cap);
                        }
                        c0 = c1; pc0 = pc1;
                        c1 = cn; pc1 = pn;
                    }
                    if (active && !slow && ca == col_cap) slow = true;          // the column is full: the list may be longer
                }
                const int n = (int)((ca - col_a) / (uint32_t)kColStride);
                n_max = max(n_max, n);

                // ---- publish the 32 lists: [n, j0, j1, ...] back to back, one reservation per warp
                const bool valid = active && !slow;
                const int words = valid ? n + 1 : 0;
                const int inc = warp_inclusive_scan(words, lane);
                const int W = __shfl_sync(kFull, inc, 31);
                const int off = inc - words;
                __syncwarp();
                if (W > 0) {
                    const unsigned long long need = a.host_out ? (unsigned long long)((W + 15) & ~15) : (unsigned long long)((W + 3) & ~3);
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(a.cursor, need);
                    base = __shfl_sync(kFull, base, 0);
                    if ((long long)(base + need) <= a.capacity) {
                        if (valid) a.list_pos[qid] = (long long)base + off;
                        // list k = column k of the hit table; (words, offset) of every list are broadcast through the (now idle) range
                        // table instead of shuffles; one coalesced store per 32 list words
                        sts_u32(tab_a, (uint32_t)words | ((uint32_t)off << 8));
                        sts_u32(tab_a + 128u, (uint32_t)inc);                                  // end offset of list k (flat expansion)
                        __syncwarp();
                        const int32_t* const out_l = a.ragged + base + lane;
                        const uint32_t id_a = slab_a + 12u;
                        const uint32_t tab_w = tab_a - (uint32_t)lane * 4u;
                        if (a.sort_lists) {
                            // ascending ids: every list is ranked in place inside its column (rank of an entry = number of entries of the
                            // list with a smaller id; ids of one list are distinct), one list at a time, entries (lane, lane + 32, ...) per lane
                            constexpr int V = (KMAX + 31) / 32;
                            for (uint32_t k = 0; k < 32u; k++) {
                                const int nk = (int)(lds_u32(tab_w + k * 4u) & 0xffu) - 1;
                                if (nk < 2) continue;
                                uint32_t ent[V];
                                int idv[V], rank[V];
#pragma unroll
                                for (int r = 0; r < V; r++) {
                                    const int e = r * 32 + lane;
                                    ent[r] = 0;
                                    idv[r] = 0x7fffffff;
                                    rank[r] = 0;
                                    if (e < nk) {
                                        ent[r] = lds_u16(col_w + (uint32_t)e * kColStride + k * 2u);
                                        idv[r] = (int)lds_u32(id_a + ent[r]);
                                    }
                                }
#pragma unroll
                                for (int rb = 0; rb < V; rb++) {
                                    const int m = min(32, nk - rb * 32);
                                    for (int j = 0; j < m; j++) {
                                        const int b = __shfl_sync(kFull, idv[rb], j);
#pragma unroll
                                        for (int r = 0; r < V; r++) rank[r] += (b < idv[r]) ? 1 : 0;
                                    }
                                }
                                __syncwarp();
#pragma unroll
                                for (int r = 0; r < V; r++)
                                    if (r * 32 + lane < nk) sts_u16(col_w + (uint32_t)rank[r] * kColStride + k * 2u, ent[r]);
                            }
                            __syncwarp();
                        }
                        if (a.host_out) {
                            // mapped host memory: PCIe wants long aligned writes, so the warp walks the FLAT word sequence of its 32 lists, a lane
                            // four consecutive words per round, and stores them as ONE 16-byte word: 512 contiguous bytes per warp instruction
                            // (measured, tools/experiments/zero_copy_store_width.py: blocks of 4 KB at scattered positions reach 52.1 GB/s with
                            // 16-byte stores per lane, 48.4 GB/s with 4-byte stores; the copy engine does 52.9 GB/s).  Word w belongs to the first
                            // list whose end offset exceeds w: binary search over the 32 end offsets for the first word, a short scan after it.
                            const int Wpad = (int)need;
                            for (int w0 = 0; w0 < Wpad; w0 += 128) {
                                const int w = w0 + 4 * lane;
                                if (w < Wpad) {
                                    const uint32_t wq = (uint32_t)min(w, W - 1);
                                    uint32_t k = (lds_u32(tab_w + 128u + 15u * 4u) <= wq) ? 16u : 0u;
                                    k += (lds_u32(tab_w + 128u + (k + 7u) * 4u) <= wq) ? 8u : 0u;
                                    k += (lds_u32(tab_w + 128u + (k + 3u) * 4u) <= wq) ? 4u : 0u;
                                    k += (lds_u32(tab_w + 128u + (k + 1u) * 4u) <= wq) ? 2u : 0u;
                                    k += (lds_u32(tab_w + 128u + k * 4u) <= wq) ? 1u : 0u;
                                    uint32_t wo = lds_u32(tab_w + k * 4u);
                                    int endk = (int)lds_u32(tab_w + 128u + k * 4u);
                                    int v[4];
#pragma unroll
                                    for (int j = 0; j < 4; j++) {
                                        const int ww = w + j;
                                        v[j] = 0;
                                        if (ww < W) {
                                            while (ww >= endk) {
                                                k++;
                                                wo = lds_u32(tab_w + k * 4u);
                                                endk = (int)lds_u32(tab_w + 128u + k * 4u);
                                            }
                                            const int u = ww - (int)(wo >> 8);
                                            v[j] = (int)(wo & 0xffu) - 1;
                                            if (u > 0) v[j] = (int)lds_u32(id_a + lds_u16(col_w + (uint32_t)(u - 1) * kColStride + k * 2u));
                                        }
                                    }
                                    asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(a.ragged + base + w), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
                                }
                            }
                        } else {
                        uint32_t colk = col_w + (uint32_t)lane * kColStride - kColStride;       // entry (lane - 1) of column k
#pragma unroll 1
                        for (int k0 = 0; k0 < 32; k0 += 4, colk += 8u) {
                            // four lists per round: all shared memory loads first, then the stores
                            uint32_t wo[4];
                            int v0[4], v1[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) wo[j] = lds_u32(tab_w + (uint32_t)(k0 + j) * 4u);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int wk = (int)(wo[j] & 0xffu);
                                v0[j] = wk - 1;
                                v1[j] = 0;
                                if (lane > 0 && lane < wk) v0[j] = (int)lds_u32(id_a + lds_u16(colk + 2u * j));
                                if (lane + 32 < wk) v1[j] = (int)lds_u32(id_a + lds_u16(colk + 2u * j + 32u * kColStride));
                            }
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int wk = (int)(wo[j] & 0xffu);
                                if (lane < wk) stg_cs_at(out_l, wo[j] >> 8, v0[j]);
                                if (lane + 32 < wk) stg_cs_at32(out_l, wo[j] >> 8, v1[j]);
                                if (KMAX > 63) {
                                    if (wk > 64) {
                                        for (int u = lane + 64; u < wk; u += 32)
                                            stg_cs_at(out_l, (wo[j] >> 8) + (uint32_t)(u - lane), (int)lds_u32(id_a + lds_u16(colk + 2u * j + (uint32_t)(u - lane) * kColStride)));
                                    }
                                }
                            }
                        }
                        }
                    } else if (lane == 0) {
                        *a.overflow = 1;
                    }
                    nb_sum += (unsigned)(W - __popc(__ballot_sync(kFull, valid)));
                }
                // ---- slow path queries, one at a time
                unsigned sm = __ballot_sync(kFull, slow);
                if (sm) {
                    const int ix_s = min(max(cx - x0, 0), ex - 1);
                    const int tq_s = (rz * kRowPitch + ry) * kTW + ix_s;                       // T entry of cell (cx - 2) of the row (cy - 2, cz - 2)
                    const int self_s = same_set ? (qp + (int)meta[44 + rr]) * 16 : -1;
                    // the warp's range table + hit columns are idle here: scratch for sorting slow-path lists
                    const uint32_t scratch_a = tab_a - (uint32_t)lane * 4u;
                    constexpr int scratch_cap = SM::kWarpBytes / 4;
                    while (sm) {
                        const int src = __ffs(sm) - 1;
                        sm &= sm - 1;
                        const float sx = __shfl_sync(kFull, q.x, src), sy = __shfl_sync(kFull, q.y, src), sz = __shfl_sync(kFull, q.z, src), sr = __shfl_sync(kFull, r2, src);
                        const int sid = __shfl_sync(kFull, qid, src);
                        if (staged)
                            brick_slow_query_staged<SYMMETRIC>(a, slab_a, buf_a + SM::kOffR2, sT + __shfl_sync(kFull, tq_s, src), sx, sy, sz, sid, sr,
                                                               __shfl_sync(kFull, self_s, src), lane, nb_sum, scratch_a, scratch_cap);
                        else
                            brick_slow_query<SYMMETRIC>(a, sx, sy, sz, sid, sr, __shfl_sync(kFull, cx, src), __shfl_sync(kFull, cy, src), __shfl_sync(kFull, cz, src), lane, nb_sum,
                                                        scratch_a, scratch_cap);
                        slow_sum++;
                    }
                }
                if (nb_sum > 0x40000000u) {
                    if (lane == 0) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
                    nb_sum = 0;
                }
                __syncwarp();
            }
            // this warp is done with the brick (and with its slab buffer)
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 16 + 8 * B);
            return true;
        };

        bool live0 = true, live1 = true;
        for (uint32_t use = 0; live0 || live1; use++) {
            if (live0) live0 = consume(std::integral_constant<int, 0>{}, use);
            if (live1) live1 = consume(std::integral_constant<int, 1>{}, use);
        }
        if (lane == 0) {
            if (nb_sum) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
            if (slow_sum) atomicAdd(a.n_slow, (unsigned long long)slow_sum);
        }
        n_max = __reduce_max_sync(kFull, n_max);
        if (lane == 0 && n_max > 0) atomicMax(a.max_list, n_max);
    }
}

}  // namespace tnsb
