// query_brick.cuh -- the fixed-radius distance query on a HALF-RADIUS grid, a lane owns a QUERY.  Default query path;
// replaces _solve_leaves / _prepare_brute_force[_simd] / _brute_force[_simd] of the reference
// (TreeNSearch.cpp:1823-1872, :2161-2399, :2400-2569).
//
// Grid: cell edge = r_max * (1 + 2^-13) / 2, linear row keys  key = (z * ny + y) * nx + x  (x fastest), prefix table
// first[key] = number of points with a smaller key (built by the bucket build, grid_build.cuh).  A neighbour of a point of cell
// (cx, cy, cz) lies in cells [cx-2, cx+2] x [cy-2, cy+2] x [cz-2, cz+2]: 25 rows, and inside a row the cells are CONSECUTIVE keys,
// i.e. one contiguous run of the sorted point array.  15.6 r^3 of candidate volume instead of the 27 r^3 of a cell = r grid, and
// per query the rows (and the x extent inside every row) that cannot hold a point within r are culled: ~75 distance tests per
// query at ~30 neighbours (2.5 tests per hit; the 27-cell stencil needs 6.5).
//
// Work decomposition (one launch per active ordered pair set_i -> set_j):
//   * brick_plan_kernel cuts the grid into bricks of <= 32 x 4 x 4 cells whose candidate SLAB (the brick plus 2 cells on every
//     side: <= 64 rows of <= 36 cells) fits the shared memory slab buffer; bricks that do not fit are split.
//   * brick_query_kernel: persistent CTAs pull bricks from a ticket counter.  Per brick one warp reads the row boundaries from
//     the prefix table and stages every slab row with ONE cp.async.bulk (global -> shared, mbarrier complete_tx) -- the rows
//     are contiguous 16-byte records (x, y, z, bits(id)); the CTA builds the brick's cell boundary table T (slab position of
//     every cell boundary) in shared memory meanwhile.
//   * a warp takes 32 consecutive queries of the brick; every LANE owns one query: it culls its 25 rows, turns them into <= 26
//     (first, last) slab ranges (its own record is cut out of its own row), and then walks ALL its ranges in one flattened loop:
//     one LDS.128 per candidate, the reference's exact arithmetic d2 = fma(dz,dz, fma(dx,dx, dy*dy)) <= r^2
//     (TreeNSearch.cpp:2477-2486 as compiled), a hit is ONE predicated 16-bit store into the lane's private
//     column.  No ballot, no popc, no shuffle in the inner loop; lanes that run out of candidates early walk a dummy range.
//   * the warp then reserves room for its 32 lists with one atomicAdd and writes them as [n, j0, j1, ...] (TreeNSearch.h:395),
//     one coalesced store per 32 list words.
// Queries with more than kMaxTot candidates, lists longer than the column, and cells too dense for any slab take a
// warp-cooperative two-pass slow path that reads the candidates from global memory.
#pragma once
#include "common.cuh"
#include "query.cuh"

namespace tnsb {

constexpr int kBX = 32, kBY = 4, kBZ = 4;                // largest brick, in cells
constexpr int kSlabRows = (kBY + 4) * (kBZ + 4);         // 64 candidate rows per brick
constexpr int kQRows = kBY * kBZ;                        // 16 query rows per brick
constexpr int kTW = kBX + 5;                             // cell boundaries per slab row (odd: spreads the rows over the banks)
constexpr int kDummy = 256;                              // dummy records (x = 3e38) behind the slab
constexpr int kTabH = 30;                                // per-lane range table: 26 ranges + 2 dummy ranges + landing entry + prefetch
constexpr int kMaxTot = 512;                             // candidates per query on the fast path
constexpr int kColStride = 68;                           // bytes between consecutive hits of one lane (34 uint16: conflict-free column reads)
constexpr uint32_t kBrickSlow = 1u << 24;                // task flag: single cell whose slab does not fit

struct BrickTask {
    int x0, y0, z0;
    uint32_t dims;      // ex | ey << 8 | ez << 16 | flags
};

struct BrickLayout {
    int off_slab, off_r2, off_T, off_rowkey, off_g0, off_rowoff, off_qs, off_qoff, off_qdelta, off_qrow, off_misc, off_warp;
    int tab_bytes, warp_bytes, total;
};

__host__ __device__ inline BrickLayout brick_layout(int slab_cap, int kmax, int n_warps, bool symmetric)
{
    BrickLayout L;
    int o = 0;
    L.off_slab = o;   o += (slab_cap + kDummy) * 16;                         // slab records, then the dummy records
    L.off_r2 = o;     o += symmetric ? (slab_cap + kDummy) * 4 : 0;          // candidate r^2 (symmetric variable radius only)
    L.off_T = o;      o += ((kSlabRows * kTW * 2 + 15) & ~15);
    L.off_rowkey = o; o += kSlabRows * 4;
    L.off_g0 = o;     o += kSlabRows * 4;
    L.off_rowoff = o; o += (kSlabRows + 4) * 4;
    L.off_qs = o;     o += kQRows * 4;
    L.off_qoff = o;   o += (kQRows + 4) * 4;
    L.off_qdelta = o; o += kQRows * 4;
    L.off_qrow = o;   o += kQRows * 4;
    L.off_misc = o;   o += 64;
    L.off_warp = o;
    L.tab_bytes = kTabH * 32 * 4;
    L.warp_bytes = (L.tab_bytes + (kmax + 1) * kColStride + 15) & ~15;
    L.total = o + n_warps * L.warp_bytes;
    return L;
}

struct BrickArgs {
    BrickGrid g;
    // searching set (set_i)
    const float4* q_pts;
    const float* q_r2;
    const uint32_t* q_first;
    // searched set (set_j)
    const float4* c_pts;
    const float* c_r2;
    const uint32_t* c_first;
    int same_set;
    int query_limit;
    float r2_fixed;
    float cull_r2;            // (largest search distance / cell)^2 with 0.1 % slack: rows farther than that are skipped
    float inv_cell_f;
    // plan
    BrickTask* tasks;
    uint32_t max_tasks;
    uint32_t* n_tasks;
    int* plan_overflow;
    uint32_t* ticket;
    int slab_cap;             // records
    int kmax;                 // hits per lane column
    // output
    int32_t* ragged;
    long long capacity;
    long long* list_pos;
    unsigned long long* cursor;
    unsigned long long* n_neighbors;
    unsigned long long* n_slow;
    int* overflow;
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

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BRICK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
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
__device__ __forceinline__ uint32_t plan_row_sum(const uint32_t* first, const BrickGrid& g, int xa, int xb, int y0, int ny_rows, int z0, int nz_rows, int lane)
{
    // sum over rows (y0 .. y0+ny_rows-1) x (z0 .. z0+nz_rows-1), clipped to the grid, of first[row + xb] - first[row + xa]
    uint32_t s = 0;
    const int n_rows = ny_rows * nz_rows;
    for (int r = lane; r < n_rows; r += 32) {
        const int y = y0 + r % ny_rows, z = z0 + r / ny_rows;
        if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) continue;
        const uint32_t key0 = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
        s += first[key0 + xb] - first[key0 + xa];
    }
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
        int sp = 0;
        if (lane == 0) {
            stack[0][0] = bx * kBX; stack[0][1] = by * kBY; stack[0][2] = bz * kBZ;
            stack[0][3] = min(kBX, g.nx - bx * kBX); stack[0][4] = min(kBY, g.ny - by * kBY); stack[0][5] = min(kBZ, g.nz - bz * kBZ);
        }
        sp = 1;
        __syncwarp();
        while (sp > 0) {
            sp--;
            const int x0 = stack[sp][0], y0 = stack[sp][1], z0 = stack[sp][2], ex = stack[sp][3], ey = stack[sp][4], ez = stack[sp][5];
            __syncwarp();
            const uint32_t nq = plan_row_sum(q_first, g, x0, x0 + ex, y0, ey, z0, ez, lane);
            if (nq == 0) continue;
            const uint32_t nc = plan_row_sum(c_first, g, max(x0 - 2, 0), min(x0 + ex + 2, g.nx), y0 - 2, ey + 4, z0 - 2, ez + 4, lane);
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
                const uint32_t id = atomicAdd(n_tasks, 1u);
                if (id < max_tasks) {
                    BrickTask t;
                    t.x0 = x0; t.y0 = y0; t.z0 = z0;
                    t.dims = (uint32_t)ex | ((uint32_t)ey << 8) | ((uint32_t)ez << 16) | flags;
                    tasks[id] = t;
                } else {
                    *plan_overflow = 1;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// slow path: one query, the whole warp, candidates from global memory, count pass + fill pass
template <bool SYMMETRIC>
__device__ __noinline__ void brick_slow_query(const BrickArgs& a, float qx, float qy, float qz, int qid, float r2, int cx, int cy, int cz, int lane, unsigned& nb_sum)
{
    const unsigned lt = lanemask_lt();
    int32_t* dst = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        int n = 0;
        for (int dz = -2; dz <= 2; dz++) {
            const int z = cz + dz;
            if (z < 0 || z >= a.g.nz) continue;
            for (int dy = -2; dy <= 2; dy++) {
                const int y = cy + dy;
                if (y < 0 || y >= a.g.ny) continue;
                const uint32_t key0 = ((uint32_t)z * (uint32_t)a.g.ny + (uint32_t)y) * (uint32_t)a.g.nx;
                const uint32_t lo = a.c_first[key0 + max(cx - 2, 0)], hi = a.c_first[key0 + min(cx + 3, a.g.nx)];
                for (uint32_t t0 = lo; t0 < hi; t0 += 32) {
                    const uint32_t t = t0 + lane;
                    bool hit = false;
                    int id = -1;
                    if (t < hi) {
                        const float4 v = a.c_pts[t];
                        id = __float_as_int(v.w);
                        const float d2 = dist2(qx, qy, qz, v.x, v.y, v.z);
                        hit = d2 <= r2;
                        if (SYMMETRIC) hit = hit || (d2 <= a.c_r2[t]);
                        if (a.same_set && id == qid) hit = false;
                    }
                    const unsigned m = __ballot_sync(kFull, hit);
                    if (pass == 1 && hit) dst[1 + n + __popc(m & lt)] = id;
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
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
template <int NWARPS, int MINBLOCKS, bool VARIABLE, bool SYMMETRIC>
__global__ void __launch_bounds__(NWARPS * 32, MINBLOCKS) brick_query_kernel(const BrickArgs a)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BrickLayout L = brick_layout(a.slab_cap, a.kmax, NWARPS, SYMMETRIC);
    const uint32_t s_base = smem_u32(s_raw);
    const uint32_t slab_a = s_base + L.off_slab;
    const uint32_t r2_a = s_base + L.off_r2;
    const uint32_t dummy_off = (uint32_t)a.slab_cap * 16u;                 // byte offset of the dummy records inside the slab
    uint16_t* const sT = reinterpret_cast<uint16_t*>(s_raw + L.off_T);
    uint32_t* const s_rowkey = reinterpret_cast<uint32_t*>(s_raw + L.off_rowkey);
    uint32_t* const s_g0 = reinterpret_cast<uint32_t*>(s_raw + L.off_g0);
    uint32_t* const s_rowoff = reinterpret_cast<uint32_t*>(s_raw + L.off_rowoff);
    uint32_t* const s_qs = reinterpret_cast<uint32_t*>(s_raw + L.off_qs);
    uint32_t* const s_qoff = reinterpret_cast<uint32_t*>(s_raw + L.off_qoff);
    int* const s_qdelta = reinterpret_cast<int*>(s_raw + L.off_qdelta);
    uint32_t* const s_qrow = reinterpret_cast<uint32_t*>(s_raw + L.off_qrow);
    uint32_t* const s_misc = reinterpret_cast<uint32_t*>(s_raw + L.off_misc);    // [0,1] mbarrier, [2] task, [3] next warp task, [4] n queries, [5] slab count
    const uint32_t bar = s_base + L.off_misc;
    const uint32_t tab_a = s_base + L.off_warp + warp * L.warp_bytes + lane * 4;  // this lane's column of the range table
    const uint32_t col_w = s_base + L.off_warp + warp * L.warp_bytes + L.tab_bytes;
    const uint32_t col_a = col_w + lane * 2;                                       // this lane's hit column
    const uint32_t col_cap = col_a + (uint32_t)a.kmax * kColStride;

    const BrickGrid g = a.g;
    const bool same_set = a.same_set != 0;
    const int query_limit = a.query_limit;
    unsigned nb_sum = 0, slow_sum = 0;

    for (int k = tid; k < kDummy; k += NWARPS * 32) {
        reinterpret_cast<float4*>(s_raw + L.off_slab)[a.slab_cap + k] = make_float4(3.0e38f, 0.0f, 0.0f, __int_as_float(-1));
        if (SYMMETRIC) reinterpret_cast<float*>(s_raw + L.off_r2)[a.slab_cap + k] = -1.0f;
    }
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t n_tasks = min(*a.n_tasks, a.max_tasks);

    for (;;) {
        if (tid == 0) s_misc[2] = atomicAdd(a.ticket, 1u);
        __syncthreads();                    // every warp is done with the previous brick's slab, tables and counters
        const uint32_t task = s_misc[2];
        if (task >= n_tasks) break;
        const BrickTask bt = a.tasks[task];
        const int x0 = bt.x0, y0 = bt.y0, z0 = bt.z0;
        const int ex = (int)(bt.dims & 0xffu), ey = (int)((bt.dims >> 8) & 0xffu), ez = (int)((bt.dims >> 16) & 0xffu);
        const bool slow_brick = (bt.dims & kBrickSlow) != 0;
        const int sy_n = ey + 4, n_rows = sy_n * (ez + 4), n_qrows = ey * ez, tw = ex + 5;

        // ---- stage: row boundaries from the prefix table, one bulk copy per slab row (warp 0); query rows (warp 1)
        if (warp == 0) {
            const int xa = max(x0 - 2, 0), xb = min(x0 + ex + 2, g.nx);
            uint32_t len[2], g0v[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int r = lane + 32 * k;
                len[k] = 0; g0v[k] = 0;
                if (r < n_rows) {
                    const int y = y0 - 2 + r % sy_n, z = z0 - 2 + r / sy_n;
                    uint32_t key0 = 0xffffffffu;
                    if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
                        key0 = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
                        g0v[k] = a.c_first[key0 + xa];
                        len[k] = a.c_first[key0 + xb] - g0v[k];
                    }
                    s_rowkey[r] = key0;
                    s_g0[r] = g0v[k];
                }
            }
            const uint32_t inc0 = (uint32_t)warp_inclusive_scan((int)len[0], lane);
            const uint32_t tot0 = __shfl_sync(kFull, inc0, 31);
            const uint32_t inc1 = (uint32_t)warp_inclusive_scan((int)len[1], lane) + tot0;
            const uint32_t total = __shfl_sync(kFull, inc1, 31);
            const uint32_t off[2] = { inc0 - len[0], inc1 - len[1] };
            if (lane < n_rows) s_rowoff[lane] = off[0];
            if (lane + 32 < n_rows) s_rowoff[lane + 32] = off[1];
            const bool fits = !slow_brick && total <= (uint32_t)a.slab_cap;
            if (lane == 0) {
                s_misc[5] = fits ? total : 0xffffffffu;
                fence_proxy_async();        // the previous brick's generic reads of the slab are ordered before the async writes
                mbar_arrive_expect_tx(bar, fits ? total * 16u : 0u);
            }
            __syncwarp();
            if (fits) {
#pragma unroll
                for (int k = 0; k < 2; k++)
                    if (len[k] > 0) bulk_g2s(slab_a + off[k] * 16u, a.c_pts + g0v[k], len[k] * 16u, bar);
            }
        } else if (warp == 1) {
            uint32_t qs = 0, cnt = 0;
            if (lane < n_qrows) {
                const int ry = lane % ey, rz = lane / ey;
                const uint32_t key0 = ((uint32_t)(z0 + rz) * (uint32_t)g.ny + (uint32_t)(y0 + ry)) * (uint32_t)g.nx;
                qs = a.q_first[key0 + x0];
                cnt = a.q_first[key0 + x0 + ex] - qs;
                s_qs[lane] = qs;
                s_qrow[lane] = (uint32_t)ry | ((uint32_t)rz << 8);
            }
            const uint32_t inc = (uint32_t)warp_inclusive_scan((int)cnt, lane);
            const uint32_t nq_all = __shfl_sync(kFull, inc, 31);
            if (lane < n_qrows) s_qoff[lane] = inc - cnt;
            if (lane == n_qrows) s_qoff[lane] = nq_all;
            if (lane == 0) { s_misc[4] = nq_all; s_misc[3] = 0u; }
        }
        __syncthreads();                    // row offsets visible
        const bool staged = s_misc[5] != 0xffffffffu;
        // ---- cell boundary table: T[row][i] = slab position of the first record of cell x0 - 2 + i of that row
        if (staged) {
            for (int e = tid; e < n_rows * tw; e += NWARPS * 32) {
                const int r = e / tw, i = e - r * tw;
                const uint32_t key0 = s_rowkey[r];
                uint32_t v = s_rowoff[r];
                if (key0 != 0xffffffffu) v += a.c_first[key0 + (uint32_t)min(max(x0 - 2 + i, 0), g.nx)] - s_g0[r];
                sT[r * kTW + i] = (uint16_t)v;
            }
            if (SYMMETRIC) {
                float* const sr2 = reinterpret_cast<float*>(s_raw + L.off_r2);
                for (int r = warp; r < n_rows; r += NWARPS) {
                    const uint32_t key0 = s_rowkey[r];
                    if (key0 == 0xffffffffu) continue;
                    const uint32_t o = s_rowoff[r], g0 = s_g0[r];
                    const uint32_t n = ((r + 1 < n_rows) ? s_rowoff[r + 1] : s_misc[5]) - o;
                    for (uint32_t k = lane; k < n; k += 32) sr2[o + k] = a.c_r2[g0 + k];
                }
            }
            if (same_set && tid < n_qrows) {
                const uint32_t qr = s_qrow[tid];
                const int srow = ((int)(qr >> 8) + 2) * sy_n + (int)(qr & 0xffu) + 2;
                s_qdelta[tid] = (int)s_rowoff[srow] - (int)s_g0[srow];
            }
        }
        __syncthreads();                    // tables visible
        mbar_wait(bar, parity);             // slab rows have landed
        parity ^= 1u;
        const int nq = (int)s_misc[4];

        // ---- warp tasks: 32 consecutive queries of the brick
        for (;;) {
            int wt = 0;
            if (lane == 0) wt = (int)atomicAdd(&s_misc[3], 1u);
            wt = __shfl_sync(kFull, wt, 0);
            if (wt * 32 >= nq) break;
            const int ci = wt * 32 + lane;
            const bool has = ci < nq;
            int rr = 0;
            for (int k = 1; k < n_qrows; k++) rr += (ci >= (int)s_qoff[k]) ? 1 : 0;
            const uint32_t qrow = s_qrow[rr];
            const int ry = (int)(qrow & 0xffu), rz = (int)(qrow >> 8);
            const int qp = has ? (int)s_qs[rr] + (ci - (int)s_qoff[rr]) : 0;
            float4 q = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0x7fffffff));
            float r2 = a.r2_fixed;
            if (has) {
                q = a.q_pts[qp];
                if (VARIABLE) r2 = a.q_r2[qp];
            }
            const int qid = __float_as_int(q.w);
            const bool active = has && qid < query_limit;
            double tx, ty, tz;
            const int cx = brick_cell(q.x, g.bottom[0], g.inv_cell, g.nx, tx);
            (void)brick_cell(q.y, g.bottom[1], g.inv_cell, g.ny, ty);
            (void)brick_cell(q.z, g.bottom[2], g.inv_cell, g.nz, tz);
            const int cy = y0 + ry, cz = z0 + rz;
            bool slow = active && (slow_brick || !staged);
            int n = 0;

            uint32_t total = 0;
            int n_ent = 0;
            if (staged) {
                // ---- this lane's candidate ranges: rows within the search distance, x extent culled per row
                const float fx = (float)(tx - (double)cx), fy = (float)(ty - (double)cy), fz = (float)(tz - (double)cz);
                float cull = a.cull_r2;
                if (VARIABLE && !SYMMETRIC) {
                    const float rc = __fmul_rn(sqrtf(r2), a.inv_cell_f);
                    cull = __fmul_rn(__fmul_rn(rc, rc), 1.002f);
                }
                const int ix = min(max(cx - x0, 0), ex - 1);
                float dd[5];
                dd[0] = fy + 1.0f; dd[1] = fy; dd[2] = 0.0f; dd[3] = 1.0f - fy; dd[4] = 2.0f - fy;
                float ddz[5];
                ddz[0] = fz + 1.0f; ddz[1] = fz; ddz[2] = 0.0f; ddz[3] = 1.0f - fz; ddz[4] = 2.0f - fz;
#pragma unroll
                for (int k = 0; k < 5; k++) { dd[k] = dd[k] * dd[k]; ddz[k] = ddz[k] * ddz[k]; }
                const int self_pos = same_set ? qp + s_qdelta[rr] : -1;
#pragma unroll
                for (int dz = 0; dz < 5; dz++) {
#pragma unroll
                    for (int dy = 0; dy < 5; dy++) {
                        const float d2yz = dd[dy] + ddz[dz];
                        const float h = sqrtf(fmaxf(cull - d2yz, 0.0f)) + 1.0e-4f;
                        const int xlo = max(__float2int_rd(fx - h), -2), xhi = min(__float2int_rd(fx + h), 2);
                        const uint16_t* trow = sT + ((rz + dz) * sy_n + (ry + dy)) * kTW + ix;
                        int lo = (int)trow[xlo + 2], hi = (int)trow[xhi + 3];
                        if (!(active && d2yz <= cull)) hi = lo;
                        if (dy == 2 && dz == 2 && self_pos >= lo && self_pos < hi) {
                            // own row: the query's own record is cut out (only the identical (set, index) is excluded, TreeNSearch.cpp:2464-2466)
                            if (self_pos > lo) {
                                sts_u32(tab_a + n_ent * 128, ((uint32_t)lo * 16u) | (((uint32_t)self_pos * 16u) << 16));
                                n_ent++;
                                total += (uint32_t)(self_pos - lo);
                            }
                            lo = self_pos + 1;
                        }
                        if (hi > lo) {
                            sts_u32(tab_a + n_ent * 128, ((uint32_t)lo * 16u) | (((uint32_t)hi * 16u) << 16));
                            n_ent++;
                            total += (uint32_t)(hi - lo);
                        }
                    }
                }
                if (total > (uint32_t)kMaxTot) { slow = true; total = 0; n_ent = 0; }
            }
            const int maxtot = (int)__reduce_max_sync(kFull, total);
            if (maxtot > 0) {
                // lanes with fewer candidates walk the dummy records (never a hit) until the longest lane is done
                int pad = maxtot - (int)total;
                while (pad > 0) {
                    const int c = min(pad, kDummy);
                    sts_u32(tab_a + n_ent * 128, dummy_off | ((dummy_off + (uint32_t)c * 16u) << 16));
                    n_ent++;
                    pad -= c;
                }
                sts_u32(tab_a + n_ent * 128, dummy_off | ((dummy_off + (uint32_t)kDummy * 16u) << 16));     // landing entry of the last advance

                // ---- the flattened candidate walk
                uint32_t w = lds_u32(tab_a);
                uint32_t p = w & 0xffffu, e = w >> 16;
                uint32_t wn = lds_u32(tab_a + 128);
                uint32_t rp = tab_a + 256;
                uint32_t ca = col_a;
                float4 c = lds_f4(slab_a + p);
#pragma unroll 2
                for (int it = 0; it < maxtot; it++) {
                    const uint32_t pc = p;
                    p += 16u;
                    if (p == e) {
                        p = wn & 0xffffu;
                        e = wn >> 16;
                        wn = lds_u32(rp);
                        rp += 128u;
                    }
                    const float4 cn = lds_f4(slab_a + p);
                    const float d2 = dist2(q.x, q.y, q.z, c.x, c.y, c.z);
                    bool hit = d2 <= r2;
                    if (SYMMETRIC) hit = hit || (d2 <= lds_f32(r2_a + (pc >> 2)));
                    if (hit) {
                        sts_u16(ca, pc);
                        ca = min(ca + (uint32_t)kColStride, col_cap);
                    }
                    c = cn;
                }
                n = (int)((ca - col_a) / (uint32_t)kColStride);
                if (active && !slow && ca == col_cap) slow = true;          // the column is full: the list may be longer
            }

            // ---- publish the 32 lists: [n, j0, j1, ...] back to back, one reservation per warp
            const bool valid = active && !slow;
            const int words = valid ? n + 1 : 0;
            const int inc = warp_inclusive_scan(words, lane);
            const int W = __shfl_sync(kFull, inc, 31);
            const int off = inc - words;
            __syncwarp();
            if (W > 0) {
                const unsigned long long need = (unsigned long long)((W + 3) & ~3);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.cursor, need);
                base = __shfl_sync(kFull, base, 0);
                if ((long long)(base + need) <= a.capacity) {
                    if (valid) a.list_pos[qid] = (long long)base + off;
                    int32_t* const out = a.ragged + base;
                    for (int k = 0; k < 32; k++) {
                        const int wk = __shfl_sync(kFull, words, k);
                        if (wk == 0) continue;
                        const int ok = __shfl_sync(kFull, off, k);
                        for (int u = lane; u < wk; u += 32) {
                            int v = wk - 1;
                            if (u > 0) v = (int)lds_u32(slab_a + lds_u16(col_w + (uint32_t)(u - 1) * kColStride + (uint32_t)k * 2u) + 12u);
                            __stcs(out + ok + u, v);
                        }
                    }
                } else if (lane == 0) {
                    *a.overflow = 1;
                }
                nb_sum += (unsigned)(W - __popc(__ballot_sync(kFull, valid)));
            }
            // ---- slow path queries, one at a time
            unsigned sm = __ballot_sync(kFull, slow);
            while (sm) {
                const int src = __ffs(sm) - 1;
                sm &= sm - 1;
                brick_slow_query<SYMMETRIC>(a, __shfl_sync(kFull, q.x, src), __shfl_sync(kFull, q.y, src), __shfl_sync(kFull, q.z, src), __shfl_sync(kFull, qid, src),
                                            __shfl_sync(kFull, r2, src), __shfl_sync(kFull, cx, src), __shfl_sync(kFull, cy, src), __shfl_sync(kFull, cz, src), lane, nb_sum);
                slow_sum++;
            }
            if (nb_sum > 0x40000000u) {
                if (lane == 0) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
                nb_sum = 0;
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        if (nb_sum) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
        if (slow_sum) atomicAdd(a.n_slow, (unsigned long long)slow_sum);
    }
}

}  // namespace tnsb
