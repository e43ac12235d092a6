// query_march.cuh -- the fixed-radius distance query as a ROW MARCH over a cell = r grid; a lane owns a CANDIDATE, hits are a bit
// matrix.  Default query path; replaces _solve_leaves / _prepare_brute_force[_simd] / _brute_force[_simd] of the reference
// (TreeNSearch.cpp:1823-1872, :2161-2399, :2400-2569).
//
// Grid: cell edge = r_max * (1 + 2^-13), linear row keys  key = (z * ny + y) * nx + x  (x fastest), prefix table
// first[key] = number of points with a smaller key, sorted records (x, y, z, bits(id)) (grid_build.cuh).  The neighbours of the
// points of cell (cx, y, z) lie in 9 rows (y-1..y+1, z-1..z+1), and in every row in the cells cx-1..cx+1, which are CONSECUTIVE
// keys: 9 contiguous runs of the sorted array, found with two loads of the prefix table each.
//
// Work decomposition (one launch per active ordered pair set_i -> set_j):
//   * march_plan_kernel lists the chunks of 8 consecutive cells of a row that hold a query point.
//   * march_query_kernel: persistent warps, no CTA-wide cooperation.  A warp pulls a chunk and marches through its cells; the run
//     bounds of the next cell and the next ticket are in flight while the current cell is processed.  Per cell:
//       FILL   the (<= NSLOT * 32) candidates of the 9 runs are packed densely over (slot, lane) and loaded ONCE into registers
//              (coalesced 16-byte loads, two slots per packed 64-bit register); their ids go to a per-warp ring in shared memory.
//       TEST   every query point of the cell is broadcast from shared memory (already duplicated into f32x2 pairs) and every lane
//              tests its candidates, two per instruction with Blackwell's packed FADD2 / FMUL2 / FFMA2, in the reference's exact
//              arithmetic d2 = fma(dz,dz, fma(dx,dx, dy*dy)) <= r^2 (TreeNSearch.cpp:2477-2486 as compiled).  The hits
//              of a (query, slot) are ONE warp ballot, stored as a word of the query's row of a bit matrix: no popc, no prefix, no
//              data-dependent store in the inner loop (5 instructions per 32 distance tests).
//     Queries accumulate over consecutive cells (they are consecutive records of the sorted array) until 32 of them are pending:
//       EXPAND a lane owns a QUERY: popc of its matrix row gives the list length, one warp scan the 32 list offsets, ONE atomicAdd
//              reserves the block of the ragged buffer, then the lane walks the set bits of its row and writes  [n, j0, j1, ...]
//              (TreeNSearch.h:395) into the warp's staging buffer; the block leaves the SM as 128-bit coalesced stores.
// Cells with more candidates than the register slots hold take a warp-cooperative two-pass path that reads from global memory.
// Self exclusion: only the identical (set, index) is excluded (TreeNSearch.cpp:2464-2466); coincident points are neighbours.
#pragma once
#include "common.cuh"
#include "query.cuh"
#include "query_brick.cuh"

namespace tnsb {

constexpr int kMarchChunk = 8;          // cells per task

// per-warp shared memory (32-bit words)
template <int NSLOT, bool SYMMETRIC>
struct MarchSmem {
    static constexpr int kOutW = NSLOT <= 8 ? 1280 : 2048;     // staging buffer of the lists of one block
    static constexpr int kRing = NSLOT * 128;                   // candidate ids of the cells of the pending queries
    static constexpr int kOffOut = 0;
    static constexpr int kOffRing = kOffOut + kOutW;
    static constexpr int kOffM = kOffRing + kRing;              // bit matrix: 32 queries x NSLOT words
    static constexpr int kOffQb = kOffM + 32 * NSLOT;           // queries of the current cell: {x, x, y, y} {z, z, r2, r2}
    static constexpr int kOffQid = kOffQb + 256;                // pending queries: id, (ring offset | own candidate number << 16)
    static constexpr int kOffMeta = kOffQid + 32;
    static constexpr int kOffRunB = kOffMeta + 32;              // the 9 runs of the cell being opened: {start, length, first candidate number, -}
    static constexpr int kOffStage = kOffRunB + 48;             // candidate records of the NEXT cell, copied asynchronously (cp.async) while the current cell is tested
    static constexpr int kOffStageR2 = kOffStage + NSLOT * 32 * 4;
    static constexpr int kWarpWords = kOffStageR2 + (SYMMETRIC ? NSLOT * 32 : 0);
    static constexpr int kWarps = (227 * 1024 / (kWarpWords * 4)) < 16 ? (227 * 1024 / (kWarpWords * 4)) : 16;
    static constexpr int kBytes = kWarps * kWarpWords * 4;
    static_assert(kOutW >= NSLOT * 32 + 4, "the staging buffer must hold one worst-case list");
    static_assert(kWarpWords % 4 == 0 && kOffM % 4 == 0 && kOffQb % 4 == 0 && kOffRing % 4 == 0 && kOffStage % 4 == 0, "16-byte alignment");
};

__device__ __forceinline__ void lds_2x64(uint32_t a, unsigned long long& v0, unsigned long long& v1)
{
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(a));
}
// cp.async with the destination given as a 32-bit shared window address
__device__ __forceinline__ void cp_async_16s(uint32_t dst, const void* gsrc) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc) : "memory"); }
__device__ __forceinline__ void cp_async_4s(uint32_t dst, const void* gsrc) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gsrc) : "memory"); }
__device__ __forceinline__ uint32_t bfind_u32(uint32_t m)
{
    uint32_t b;
    asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(m));
    return b;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void stg_cs_u4(int32_t* p, const uint4& v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// chunks of kMarchChunk consecutive cells of a row that hold at least one point of the searching set, in key order (a warp looks at
// 32 consecutive chunks and reserves their task slots with one atomicAdd)
__global__ void __launch_bounds__(256) march_plan_kernel(const BrickGrid g, const uint32_t* __restrict__ q_first, BrickTask* __restrict__ tasks, uint32_t max_tasks,
                                                         uint32_t* __restrict__ n_tasks, int* __restrict__ plan_overflow)
{
    const int lane = threadIdx.x & 31;
    const int chunks = ceil_div(g.nx, kMarchChunk);
    const long long total = (long long)chunks * g.ny * g.nz;
    const unsigned lt = lanemask_lt();
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < total; base += (long long)gridDim.x * blockDim.x) {
        const long long t = base + lane;
        bool has = false;
        int x0 = 0, row = 0, nc = 0;
        if (t < total) {
            row = (int)(t / chunks);
            x0 = (int)(t % chunks) * kMarchChunk;
            nc = min(kMarchChunk, g.nx - x0);
            const uint32_t rk = (uint32_t)row * (uint32_t)g.nx + (uint32_t)x0;
            has = q_first[rk + nc] > q_first[rk];
        }
        const unsigned m = __ballot_sync(kFull, has);
        if (m == 0u) continue;
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(n_tasks, (uint32_t)__popc(m));
        b = __shfl_sync(kFull, b, 0);
        if (has) {
            const uint32_t id = b + (uint32_t)__popc(m & lt);
            if (id < max_tasks) {
                BrickTask bt;
                bt.x0 = x0; bt.y0 = row % g.ny; bt.z0 = row / g.ny;
                bt.dims = (uint32_t)nc;
                tasks[id] = bt;
            } else {
                *plan_overflow = 1;
            }
        }
    }
}

// ascending sort of n <= 32 * V ints held as element e = r * 32 + lane in v[r] (padding: INT_MAX): bitonic network in registers
template <int V>
__device__ __forceinline__ void warp_bitonic_regs(int (&v)[V], int lane)
{
#pragma unroll
    for (int k = 2; k <= 32 * V; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 1; j >>= 1) {
            int o[V];
#pragma unroll
            for (int r = 0; r < V; r++) o[r] = v[r];
#pragma unroll
            for (int r = 0; r < V; r++) {
                const int e = r * 32 + lane;
                int other;
                if (j >= 32) other = o[r ^ (j >> 5)];
                else other = __shfl_xor_sync(kFull, o[r], j);
                const bool up = (e & k) == 0;
                const bool lower = (e & j) == 0;
                const int lo = min(o[r], other), hi = max(o[r], other);
                v[r] = (lower == up) ? lo : hi;
            }
        }
    }
}

template <int NSLOT, bool VARIABLE, bool SYMMETRIC>
__global__ void __launch_bounds__(MarchSmem<NSLOT, SYMMETRIC>::kWarps * 32, 1) march_query_kernel(const BrickArgs a)
{
    typedef MarchSmem<NSLOT, SYMMETRIC> SM;
    static_assert(NSLOT % 4 == 0, "matrix rows are stored as 128-bit words");
    constexpr int NPMAX = NSLOT / 2;
    extern __shared__ __align__(16) uint32_t s_march[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* const wm = s_march + warp * SM::kWarpWords;
    uint32_t wa = smem_u32(wm);
    asm volatile("" : "+r"(wa));
    const uint32_t out_a = wa + SM::kOffOut * 4, ring_a = wa + SM::kOffRing * 4, m_a = wa + SM::kOffM * 4, qb_a = wa + SM::kOffQb * 4;
    const uint32_t qid_a = wa + SM::kOffQid * 4, meta_a = wa + SM::kOffMeta * 4, stage_a = wa + SM::kOffStage * 4, stage_r2_a = wa + SM::kOffStageR2 * 4;
    const uint32_t runs_a = wa + SM::kOffRunB * 4;
    const BrickGrid g = a.g;
    const unsigned lt = lanemask_lt();
    const float r2_fixed = a.r2_fixed;
    const int query_limit = a.query_limit;
    const bool same_set = a.same_set != 0;
    const uint32_t n_tasks = min(*a.n_tasks, a.max_tasks);

    int nqb = 0;            // pending queries (rows of the bit matrix in use)
    int ring_used = 0;      // candidate ids in the ring
    unsigned nb_sum = 0, slow_sum = 0, over8_sum = 0;
    int n_max = 0, t_max = 0;

    // The march is ONE flat loop over (flush | open the next cell | take it | fill | one group of queries): every stage exists once in
    // the code.  The NEXT cell is opened -- run bounds, candidate copies into the staging buffer, first queries -- right after the
    // current cell's candidates have moved from the staging buffer into registers, so its global memory latency hides under the
    // current cell's tests.
    bool have_cell = false, finished = false, must_flush = false, need_open = true, opened_next = false, filled = false;
    // chunk state
    bool have_task = false;
    uint32_t t_cur = 0, t_next = 0;
    int x1 = 0, ty = 0, tz = 0, cx = 0;     // cx: the cell whose run bounds are in flight in (lo, hi)
    bool row_ok = false;
    const uint32_t* tab = a.c.first;
    uint32_t rk = 0, lo = 0, hi = 0;
    // the opened (next) cell
    bool p_valid = false;
    int p_x = 0, p_y = 0, p_z = 0, p_qs = 0, p_qe = 0, p_T = 0, p_lo_own = 0, p_pre_own = 0;
    float4 p_qv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float p_qr2 = 0.0f;
    // the current cell
    int cell_x = 0, cell_y = 0, cell_z = 0, qs = 0, qe = 0, q0 = 0, T = 0, lo_own = 0, pre_own = 0, cell_ring = 0;
    float4 c_qv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float c_qr2 = 0.0f;
    f32x2 px[NPMAX], py[NPMAX], pz[NPMAX];
    float pr2[SYMMETRIC ? NSLOT : 1];

    // lanes 0..8: the 9 candidate rows (run bounds from the searched set's table); lane 9: the query row (searching set's table)
    auto load_bounds = [&](int x) {
        lo = 0; hi = 0;
        if (row_ok) {
            lo = tab[rk + (uint32_t)max(x - 1, 0)];
            hi = tab[rk + (uint32_t)min(x + 2, g.nx)];
        } else if (lane == 9) {
            lo = tab[rk + (uint32_t)x];
            hi = tab[rk + (uint32_t)x + 1u];
        }
    };

    if (lane == 0) t_cur = atomicAdd(a.ticket, 1u);
    t_cur = __shfl_sync(kFull, t_cur, 0);

    for (;;) {
        if (must_flush) {
            // ================= EXPAND + publish the pending queries =================
            __syncwarp();
            if (nqb > 0) {
                const bool inb = lane < nqb;
                int qid = 0x7fffffff;
                uint32_t meta = 0xffff0000u;
                if (inb) {
                    qid = (int)lds_u32(qid_a + lane * 4);
                    meta = lds_u32(meta_a + lane * 4);
                }
                const bool active = inb && qid < query_limit;
                const uint32_t fs = meta >> 16;           // the query's own candidate number (0xffff: it is not a candidate)
                const uint32_t ra = ring_a + (meta & 0xffffu) * 4u;
                const uint32_t row_a = m_a + (uint32_t)(lane * NSLOT) * 4u;
                if (active && fs != 0xffffu) {
                    const uint32_t wa_ = row_a + (fs >> 5) * 4u;
                    sts_u32(wa_, lds_u32(wa_) & ~(1u << (fs & 31u)));
                }
                uint32_t mw[NSLOT];
                int n = 0;
#pragma unroll
                for (int w4 = 0; w4 < NSLOT / 4; w4++) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (active) v = lds_u4(row_a + (uint32_t)w4 * 16u);
                    mw[w4 * 4 + 0] = v.x; mw[w4 * 4 + 1] = v.y; mw[w4 * 4 + 2] = v.z; mw[w4 * 4 + 3] = v.w;
                }
#pragma unroll
                for (int w = 0; w < NSLOT; w++) n += __popc(mw[w]);
                n_max = max(n_max, n);
                const int words = active ? n + 1 : 0;
                const int inc = warp_inclusive_scan(words, lane);
                const int W = __shfl_sync(kFull, inc, 31);
                int start = 0, jb = 0;
                while (start < W) {
                    // the longest run of lists [jb, je) that fits the staging buffer
                    const unsigned fits = __ballot_sync(kFull, lane >= jb && inc - start <= SM::kOutW);
                    const int je = jb + __popc(fits);
                    const int Wsub = __shfl_sync(kFull, inc, je - 1) - start;
                    const bool mine = lane >= jb && lane < je && words > 0;
                    if (Wsub > 0) {
                        const unsigned long long need = a.host_out ? (unsigned long long)((Wsub + 15) & ~15) : (unsigned long long)((Wsub + 3) & ~3);
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(a.cursor, need);
                        base = __shfl_sync(kFull, base, 0);
                        if ((long long)(base + need) <= a.capacity) {
                            const int o = inc - words - start;
                            if (mine) {
                                a.list_pos[qid] = (long long)base + o;
                                uint32_t p = out_a + (uint32_t)o * 4u;
                                sts_u32(p, (uint32_t)n);
                                p += 4u;
#pragma unroll
                                for (int w = 0; w < NSLOT; w++) {
                                    uint32_t m = mw[w];
                                    const uint32_t rw = ra + (uint32_t)w * 128u;
                                    while (m) {
                                        const uint32_t b = bfind_u32(m);
                                        m ^= 1u << b;
                                        sts_u32(p, lds_u32(rw + b * 4u));
                                        p += 4u;
                                    }
                                }
                            }
                            __syncwarp();
                            if (a.sort_lists) {
                                // ascending ids (the reference's order): every list is sorted in the staging buffer
                                for (int k = jb; k < je; k++) {
                                    const int nk = __shfl_sync(kFull, n, k);
                                    const int wk = __shfl_sync(kFull, words, k);
                                    const uint32_t la = out_a + (uint32_t)(__shfl_sync(kFull, o, k) + 1) * 4u;
                                    if (wk == 0 || nk < 2) continue;
                                    if (nk <= 32) {
                                        int v[1] = { lane < nk ? (int)lds_u32(la + lane * 4u) : 0x7fffffff };
                                        warp_bitonic_regs<1>(v, lane);
                                        if (lane < nk) sts_u32(la + lane * 4u, (uint32_t)v[0]);
                                    } else if (nk <= 64) {
                                        int v[2];
#pragma unroll
                                        for (int r = 0; r < 2; r++) v[r] = r * 32 + lane < nk ? (int)lds_u32(la + (uint32_t)(r * 32 + lane) * 4u) : 0x7fffffff;
                                        warp_bitonic_regs<2>(v, lane);
#pragma unroll
                                        for (int r = 0; r < 2; r++)
                                            if (r * 32 + lane < nk) sts_u32(la + (uint32_t)(r * 32 + lane) * 4u, (uint32_t)v[r]);
                                    } else {
                                        warp_bitonic_sort(nk, lane, [&](int i) { return (int)lds_u32(la + (uint32_t)i * 4u); }, [&](int i, int x) { sts_u32(la + (uint32_t)i * 4u, (uint32_t)x); });
                                    }
                                    __syncwarp();
                                }
                            }
                            int32_t* const dst = a.ragged + base;
                            for (int t = lane * 4; t < Wsub; t += 128) stg_cs_u4(dst + t, lds_u4(out_a + (uint32_t)t * 4u));
                            __syncwarp();
                            nb_sum += (unsigned)(Wsub - __popc(__ballot_sync(kFull, mine)));
                        } else if (lane == 0) {
                            *a.overflow = 1;
                        }
                    }
                    start += Wsub;
                    jb = je;
                }
            }
            nqb = 0;
            if (!(have_cell && filled)) ring_used = 0;       // a cell with queries still to come keeps its ids where they are
            if (nb_sum > 0x40000000u) {
                if (lane == 0) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
                nb_sum = 0;
            }
            must_flush = false;
            if (finished) break;
        }

        if (need_open) {
            // ================= OPEN the next cell that holds queries: its 9 runs, its candidates on their way into the staging buffer,
            // its first 32 queries on their way into registers (next chunk from the ticket counter, requested one chunk ahead) =================
            need_open = false;
            opened_next = true;
            p_valid = false;
            for (;;) {
                if (!have_task) {
                    if (t_cur >= n_tasks) break;
                    if (lane == 0) t_next = atomicAdd(a.ticket, 1u);
                    const BrickTask bt = a.tasks[t_cur];
                    cx = bt.x0; ty = bt.y0; tz = bt.z0;
                    x1 = bt.x0 + (int)bt.dims;
                    const int yy = ty + lane % 3 - 1, zz = tz + lane / 3 - 1;
                    row_ok = lane < 9 && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz;
                    tab = a.c.first;
                    rk = ((uint32_t)zz * (uint32_t)g.ny + (uint32_t)yy) * (uint32_t)g.nx;
                    if (lane == 9) {
                        tab = a.q.first;
                        rk = ((uint32_t)tz * (uint32_t)g.ny + (uint32_t)ty) * (uint32_t)g.nx;
                    }
                    load_bounds(cx);
                    have_task = true;
                }
                const uint32_t cur_lo = lo, cur_hi = hi;
                p_x = cx; p_y = ty; p_z = tz;
                cx++;
                if (cx < x1) {
                    load_bounds(cx);
                } else {
                    have_task = false;
                    t_cur = __shfl_sync(kFull, t_next, 0);
                }
                p_qs = (int)__shfl_sync(kFull, cur_lo, 9);
                p_qe = (int)__shfl_sync(kFull, cur_hi, 9);
                if (p_qe <= p_qs) continue;
                const int len = lane < 9 ? (int)(cur_hi - cur_lo) : 0;
                int inc = len;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const int t = __shfl_up_sync(kFull, inc, o);
                    if (lane >= o) inc += t;
                }
                p_T = __shfl_sync(kFull, inc, 8);
                const int pre = inc - len;
                t_max = max(t_max, p_T);
                if (p_T > 256) over8_sum += (unsigned)(p_qe - p_qs);
                p_lo_own = __shfl_sync(kFull, (int)cur_lo, 4);
                p_pre_own = __shfl_sync(kFull, pre, 4);
                if (lane < p_qe - p_qs) {
                    p_qv = a.q.pts[p_qs + lane];
                    if (VARIABLE) p_qr2 = a.q.r2[p_qs + lane];
                }
                if (p_T <= NSLOT * 32) {
                    // record (first candidate number of the run + i) of the staging buffer <- record i of the run: 16-byte asynchronous copies,
                    // one run per step (its descriptor is broadcast from shared memory), a lane per record
                    if (lane < 9) sts_u4(runs_a + (uint32_t)lane * 16u, cur_lo, (uint32_t)len, (uint32_t)pre, 0u);
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 9; r++) {
                        const uint4 d = lds_u4(runs_a + (uint32_t)r * 16u);
                        if ((uint32_t)lane < d.y) {
                            cp_async_16s(stage_a + (d.z + (uint32_t)lane) * 16u, a.c.pts + (d.x + (uint32_t)lane));
                            if (SYMMETRIC) cp_async_4s(stage_r2_a + (d.z + (uint32_t)lane) * 4u, a.c.r2 + (d.x + (uint32_t)lane));
                        }
                        if (d.y > 32u) {
                            // runs longer than a warp (dense rows): the rest, one warp-wide step at a time
#pragma unroll 1
                            for (uint32_t i = (uint32_t)lane + 32u; i < d.y; i += 32u) {
                                cp_async_16s(stage_a + (d.z + i) * 16u, a.c.pts + (d.x + i));
                                if (SYMMETRIC) cp_async_4s(stage_r2_a + (d.z + i) * 4u, a.c.r2 + (d.x + i));
                            }
                        }
                    }
                    __syncwarp();      // the run table may be rewritten by the next open
                }
                cp_async_commit();
                p_valid = true;
                break;
            }
        }

        if (!have_cell) {
            // ================= take the opened cell =================
            if (!p_valid) {
                finished = true;
                must_flush = true;
                continue;
            }
            cell_x = p_x; cell_y = p_y; cell_z = p_z;
            qs = p_qs; qe = p_qe; T = p_T; lo_own = p_lo_own; pre_own = p_pre_own;
            c_qv = p_qv; c_qr2 = p_qr2;
            q0 = qs;
            filled = false;
            opened_next = false;
            have_cell = true;
        }

        if (T > NSLOT * 32) {
            // ================= dense neighbourhood: one query at a time, the whole warp, two passes over the candidates in global memory
            if (nqb > 0) {
                must_flush = true;
                continue;
            }
            if (!opened_next) {
                need_open = true;
                continue;
            }
            for (int qi = qs; qi < qe; qi++) {
                const float4 qv = a.q.pts[qi];
                const int qid = __float_as_int(qv.w);
                if (qid >= query_limit) continue;
                const float r2 = VARIABLE ? a.q.r2[qi] : r2_fixed;
                brick_slow_query<SYMMETRIC, 1>(a, qv.x, qv.y, qv.z, qid, r2, cell_x, cell_y, cell_z, lane, nb_sum, out_a, SM::kOutW);
                slow_sum++;
                if (nb_sum > 0x40000000u) {
                    if (lane == 0) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
                    nb_sum = 0;
                }
            }
            __syncwarp();
            have_cell = false;
            continue;
        }

        const int Qg = min(32, qe - q0);
        if (nqb + Qg > 32 || (!filled && ring_used + T > SM::kRing)) {
            must_flush = true;
            continue;
        }
        if (!filled) {
            // ================= FILL: staging buffer -> candidate registers (two slots per packed 64-bit register), ids -> ring =================
            cp_async_wait_all();
            __syncwarp();
            cell_ring = ring_used;
            const uint32_t ring_w = ring_a + (uint32_t)(cell_ring + lane) * 4u;
            const uint32_t st_l = stage_a + (uint32_t)lane * 16u;
            const int npairs_fill = max((T + 63) >> 6, 1);
#pragma unroll
            for (int j = 0; j < NPMAX; j++) {
                if (j < npairs_fill) {
                    // a lane without a candidate holds a point at x = 3e38: d2 = inf or NaN, never a hit (and r2 = -1 for the symmetric test);
                    // the stale record it reads (and the stale id it stores) is never looked at
                    float4 v0 = lds_f4(st_l + (uint32_t)(2 * j * 32) * 16u);
                    float4 v1 = lds_f4(st_l + (uint32_t)((2 * j + 1) * 32) * 16u);
                    const bool ok0 = 2 * j * 32 + lane < T, ok1 = (2 * j + 1) * 32 + lane < T;
                    if (ok0) sts_u32(ring_w + (uint32_t)(2 * j * 32) * 4u, (uint32_t)__float_as_int(v0.w));
                    if (ok1) sts_u32(ring_w + (uint32_t)((2 * j + 1) * 32) * 4u, (uint32_t)__float_as_int(v1.w));
                    v0.x = ok0 ? v0.x : 3.0e38f;
                    v1.x = ok1 ? v1.x : 3.0e38f;
                    px[j] = settle2(pack2(v0.x, v1.x));
                    py[j] = settle2(pack2(v0.y, v1.y));
                    pz[j] = settle2(pack2(v0.z, v1.z));
                    if (SYMMETRIC) {
                        const float w0 = lds_f32(stage_r2_a + (uint32_t)(2 * j * 32 + lane) * 4u), w1 = lds_f32(stage_r2_a + (uint32_t)((2 * j + 1) * 32 + lane) * 4u);
                        pr2[SYMMETRIC ? 2 * j : 0] = ok0 ? w0 : -1.0f;
                        pr2[SYMMETRIC ? 2 * j + 1 : 0] = ok1 ? w1 : -1.0f;
                    }
                }
            }
            __syncwarp();          // the staging buffer is free for the next cell
            ring_used += T;
            filled = true;
            need_open = true;
            continue;
        }
        // ================= the queries of the group: broadcast layout for the test loop, (id, ring offset, own candidate number) for the expansion
        if (lane < Qg) {
            float4 qv = c_qv;
            float r2 = VARIABLE ? c_qr2 : r2_fixed;
            if (q0 != qs) {
                qv = a.q.pts[q0 + lane];
                if (VARIABLE) r2 = a.q.r2[q0 + lane];
            }
            const uint32_t qa = qb_a + (uint32_t)lane * 32u;
            sts_u4(qa, __float_as_uint(qv.x), __float_as_uint(qv.x), __float_as_uint(qv.y), __float_as_uint(qv.y));
            sts_u4(qa + 16u, __float_as_uint(qv.z), __float_as_uint(qv.z), __float_as_uint(r2), __float_as_uint(r2));
            const uint32_t fs = same_set ? (uint32_t)(pre_own + (q0 + lane - lo_own)) : 0xffffu;
            sts_u32(qid_a + (uint32_t)(nqb + lane) * 4u, (uint32_t)__float_as_int(qv.w));
            sts_u32(meta_a + (uint32_t)(nqb + lane) * 4u, (uint32_t)cell_ring | (fs << 16));
        }
        __syncwarp();
        // ================= TEST: every query of the group against the candidate registers; row (nqb + k) of the bit matrix receives the ballots
        auto run_queries = [&](auto np_tag) {
            constexpr int NP = decltype(np_tag)::value;
            uint32_t qa = qb_a;
            uint32_t rowa = m_a + (uint32_t)(nqb * NSLOT) * 4u;
#pragma unroll 2
            for (int k = 0; k < Qg; k++, qa += 32u, rowa += NSLOT * 4u) {
                f32x2 qx, qy, qz, qr;
                lds_2x64(qa, qx, qy);
                lds_2x64(qa + 16u, qz, qr);
                float r2, r2_hi;
                unpack2(qr, r2, r2_hi);
                uint32_t mm[NSLOT];
#pragma unroll
                for (int w = 0; w < NSLOT; w++) mm[w] = 0u;
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const f32x2 dx = sub2(qx, px[j]);
                    const f32x2 dy = sub2(qy, py[j]);
                    const f32x2 dz = sub2(qz, pz[j]);
                    const f32x2 d2p = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
                    float d2[2];
                    unpack2(d2p, d2[0], d2[1]);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        bool hit = d2[h] <= r2;
                        if (SYMMETRIC) hit = hit || (d2[h] <= pr2[SYMMETRIC ? 2 * j + h : 0]);
                        mm[2 * j + h] = __ballot_sync(kFull, hit);
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int w4 = 0; w4 < NSLOT / 4; w4++) sts_u4(rowa + w4 * 16u, mm[w4 * 4], mm[w4 * 4 + 1], mm[w4 * 4 + 2], mm[w4 * 4 + 3]);
                }
            }
        };
        const int npairs = max((T + 63) >> 6, 1);
        if constexpr (NSLOT == 8) {
            switch (npairs) {
            case 1: run_queries(std::integral_constant<int, 1>{}); break;
            case 2: run_queries(std::integral_constant<int, 2>{}); break;
            case 3: run_queries(std::integral_constant<int, 3>{}); break;
            default: run_queries(std::integral_constant<int, 4>{}); break;
            }
        } else {
            switch (npairs) {
            case 1: case 2: run_queries(std::integral_constant<int, 2>{}); break;
            case 3: case 4: run_queries(std::integral_constant<int, 4>{}); break;
            case 5: run_queries(std::integral_constant<int, 5>{}); break;
            case 6: run_queries(std::integral_constant<int, 6>{}); break;
            case 7: run_queries(std::integral_constant<int, 7>{}); break;
            default: run_queries(std::integral_constant<int, 8>{}); break;
            }
        }
        __syncwarp();          // the group's broadcast records may be overwritten by the next group
        nqb += Qg;
        q0 += 32;
        if (q0 >= qe) have_cell = false;
    }

    if (lane == 0) {
        if (nb_sum) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
        if (slow_sum) atomicAdd(a.n_slow, (unsigned long long)slow_sum);
        if (over8_sum) atomicAdd(a.n_over8, (unsigned long long)over8_sum);
    }
    n_max = __reduce_max_sync(kFull, n_max);
    if (lane == 0 && n_max > 0) atomicMax(a.max_list, n_max);
    if (lane == 0 && t_max > 0) atomicMax(a.max_cand, t_max);
}

}  // namespace tnsb
