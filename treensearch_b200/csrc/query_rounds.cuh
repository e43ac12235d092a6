// query_rounds.cuh -- second generation of the 27-cell fixed-radius query ("round" kernel).  Same contract and arguments as
// query_kernel (query.cuh); replaces _solve_leaves / _prepare_brute_force[_simd] / _brute_force[_simd] of the reference
// (TreeNSearch.cpp:1823-1872, :2161-2399, :2400-2569).
//
// Why a second kernel: query_kernel keeps the candidates in registers and gives every LANE a candidate, so each
// (query, 32 candidates) step needs a ballot, two popcounts and an address computation to compact the hits -- ncu shows it
// bound by instruction issue (~190 warp instructions per query at 10M points), not by memory.  Here the roles are swapped:
//
//   * every LANE owns a QUERY.  A warp forms a "round" of up to 32 consecutive (Morton ordered) query points taken from up to
//     NT consecutive occupied cells;
//   * for each of those cells the 27 neighbour runs are looked up (dense Morton table or hash) and the candidates are staged
//     ONCE into a shared-memory tile in structure-of-arrays form X[t], Y[t], Z[t], ID[t] (coalesced 16-byte loads, conflict
//     free 4-byte stores);
//   * inner loop: a lane walks the tile of ITS cell two candidates at a time -- three LDS.64 give (x0,x1) (y0,y1) (z0,z1) as
//     packed pairs, six packed FADD2/FMUL2/FFMA2 give both distances in the reference's exact arithmetic
//     d2 = fma(dz,dz, fma(dx,dx, dy*dy)) (TreeNSearch.cpp:2477-2486 as compiled), and a hit costs one predicated
//     byte store plus one predicated pointer bump into the lane's PRIVATE hit list.  No ballots, no popcounts, no shuffles, no
//     divergence: lanes of different cells read different tiles (bank-staggered, one wavefront per LDS);
//   * after the loop the private lists (one byte per hit: the candidate's pair number) are expanded to ids and transposed
//     into a contiguous staging buffer in the reference's  [n, j0, j1, ...]  layout (TreeNSearch.h:395) -- the staging buffer
//     aliases the X/Y/Z tiles, which are dead by then -- and flushed with ONE atomicAdd and fully coalesced 128-bit stores.
//
// Self exclusion: only the identical (set, index) is excluded (TreeNSearch.cpp:2464-2466); the query always finds itself
// (d2 = 0) and the entry is dropped while the list is expanded.  Coincident points are neighbours.
#pragma once
#include "query.cuh"

namespace tnsb {

constexpr int kRoundCells = 32;      // cells per ticket (one key / start per lane)

// per-warp shared memory layout (in 4-byte words)
template <int NT, bool SYMMETRIC>
struct RLayout {
    static constexpr int kTileCap = 1024 / NT;            // candidates per tile (cells with more take the slow path)
    static constexpr int kArr = kTileCap + 2;             // words per coordinate array; the 2 extra words stagger the banks of the NT tiles
    static constexpr int kStageInts = NT * 3 * kArr;      // X/Y/Z of all tiles; doubles as the output staging buffer
    static constexpr int kOffR2 = kStageInts;             // candidate r^2 (symmetric variable radius only)
    static constexpr int kOffId = kOffR2 + (SYMMETRIC ? NT * kArr : 0);
    static constexpr int kCapSub = 96;                    // rows of each private sub-list (even / odd candidates)
    static constexpr int kOffLists = kOffId + NT * kTileCap;
    static constexpr int kOffRuns = kOffLists + 2 * kCapSub * 32 / 4;
    static constexpr int kWarpInts = (kOffRuns + 32 + 3) & ~3;
    static constexpr int kWarps = SYMMETRIC ? 8 : 10;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kBytes = kWarps * kWarpInts * 4;
    static_assert(kArr % 2 == 0 && kOffR2 % 2 == 0, "8-byte alignment of the packed coordinate pairs");
    static_assert(kTileCap / 2 <= 256, "a pair number must fit one byte");
    static_assert(kBytes <= 227 * 1024, "shared memory per SM");
};

template <typename Key, bool DENSE>
struct NeighborLookup {
    Key key;
    uint32_t slot;
    bool valid;
    typename HashSlot<Key>::Raw e;
    uint2 d;
};

// lanes 0..26 own one neighbour cell each; the load is issued here and consumed by lookup_resolve
template <typename Key, bool DENSE>
__device__ __forceinline__ void lookup_issue(NeighborLookup<Key, DENSE>& L, const QueryArgs<Key>& a, Key cell_key, int lane)
{
    int l = lane;
    asm volatile("" : "+r"(l));       // keeps the per-lane Morton constants from being hoisted into registers for the whole kernel
    const int ox = l % 3 - 1, oy = (l / 3) % 3 - 1, oz = l / 9 - 1;
    L.valid = l < 27;
    L.key = morton_neighbor<Key>(cell_key, ox, oy, oz, a.key_mask, L.valid);
    L.d = make_uint2(0u, 0u);
    L.slot = 0;
    if (DENSE) {
        if (L.valid) L.d = __ldg(a.dense + L.key);
    } else {
        L.slot = Morton<Key>::hash(L.key) >> (32 - a.hash_log2);
        if (L.valid) L.e = HashSlot<Key>::load(a.htable, L.slot);
    }
}

template <typename Key, bool DENSE>
__device__ __forceinline__ void lookup_resolve(const NeighborLookup<Key, DENSE>& L, const QueryArgs<Key>& a, int& rs, int& rc)
{
    rs = 0;
    rc = 0;
    if (DENSE) {
        rs = (int)L.d.x;
        rc = (int)(L.d.y - L.d.x);
    } else if (L.valid) {
        const uint32_t hmask = (1u << a.hash_log2) - 1u;
        typename HashSlot<Key>::Raw e = L.e;
        uint32_t slot = L.slot;
        for (;;) {
            if (HashSlot<Key>::matches(e, L.key)) { rs = HashSlot<Key>::start(e); rc = HashSlot<Key>::count(e); break; }
            if (HashSlot<Key>::is_empty(e)) break;
            slot = (slot + 1) & hmask;
            e = HashSlot<Key>::load(a.htable, slot);
        }
    }
}

// Turns the 27 (start, count) runs into one dense candidate numbering: returns T (candidates of the cell), leaves this lane's
// first candidate number in `pre` and writes run_base[k] = start - pre of the k-th non-empty run for candidate_pos().
__device__ __forceinline__ int build_runs(int rs, int rc, int* run_base, int lane, unsigned lt, int& pre)
{
    const int inc = warp_inclusive_scan(rc, lane);
    const int T = __shfl_sync(kFull, inc, 31);
    pre = inc - rc;
    const unsigned nonempty = __ballot_sync(kFull, rc > 0);
    if (rc > 0) run_base[__popc(nonempty & lt)] = rs - pre;
    __syncwarp();
    return T;
}

__device__ __forceinline__ unsigned smem_addr_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// one-byte store into the lane's private hit list.  Deliberately without a "memory" clobber: the list is only read back after
// a __syncwarp(), and the candidate loads of the following pairs must stay free to be scheduled above it.
__device__ __forceinline__ void sts_u8(unsigned addr, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v)); }

template <int N>
__device__ __forceinline__ int pick(const int (&v)[N], int s)
{
    int r = v[0];
#pragma unroll
    for (int i = 1; i < N; i++)
        if (s == i) r = v[i];
    return r;
}

// Slow path for the queries [qb, qe) of one cell whose neighbourhood does not fit a tile (or whose round overflowed a private
// list): two sweeps per query over the candidates through L1/L2, lists written straight to the ragged buffer.
template <typename Key, bool VARIABLE, bool SYMMETRIC, bool DENSE>
__device__ __forceinline__ unsigned slow_queries(const QueryArgs<Key>& a, Key cell_key, int qb, int qe, int cell_qb, int* run_base, int lane, unsigned lt)
{
    NeighborLookup<Key, DENSE> L;
    lookup_issue(L, a, cell_key, lane);
    int rs, rc, pre;
    lookup_resolve(L, a, rs, rc);
    __syncwarp();
    const int T = build_runs(rs, rc, run_base, lane, lt, pre);
    const int self_pre = __shfl_sync(kFull, pre, 13);       // lane 13 = offset (0,0,0): the cell itself when same_set
    unsigned found = 0;
    for (int qi = qb; qi < qe; qi++) {
        const float4 qv = a.q_pts[qi];
        const int qidx = __float_as_int(qv.w);
        if (qidx >= a.query_limit) continue;
        const float r2 = VARIABLE ? a.q_r2[qi] : a.r2_fixed;
        const int ts = a.same_set ? self_pre + (qi - cell_qb) : -1;
        int n = 0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            const int pos = candidate_pos(pre, rc, run_base, t0, lane);
            bool h = false;
            if (t < T && t != ts) {
                const float4 v = a.c_pts[pos];
                const float d2 = dist2(qv.x, qv.y, qv.z, v.x, v.y, v.z);
                h = d2 <= r2;
                if (SYMMETRIC) h = h || (d2 <= a.c_r2[pos]);
            }
            n += __popc(__ballot_sync(kFull, h));
        }
        const unsigned long long need = (unsigned long long)((n + 1 + 3) & ~3);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.cursor, need);
        base = __shfl_sync(kFull, base, 0);
        if ((long long)(base + need) > a.capacity) {
            if (lane == 0) *a.overflow = 1;
            continue;
        }
        if (lane == 0) {
            a.ragged[base] = n;
            a.list_pos[qidx] = (long long)base;
        }
        int32_t* dst = a.ragged + base;
        int p = 1;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            const int pos = candidate_pos(pre, rc, run_base, t0, lane);
            bool h = false;
            int id = -1;
            if (t < T && t != ts) {
                const float4 v = a.c_pts[pos];
                const float d2 = dist2(qv.x, qv.y, qv.z, v.x, v.y, v.z);
                h = d2 <= r2;
                if (SYMMETRIC) h = h || (d2 <= a.c_r2[pos]);
                id = __float_as_int(v.w);
            }
            const unsigned m = __ballot_sync(kFull, h);
            if (h) dst[p + __popc(m & lt)] = id;
            p += __popc(m);
        }
        found += (unsigned)n;
    }
    __syncwarp();
    return found;
}

template <typename Key, int NT, bool VARIABLE, bool SYMMETRIC, bool DENSE>
__global__ void __launch_bounds__(RLayout<NT, SYMMETRIC>::kThreads, 1) query_rounds_kernel(const QueryArgs<Key> a)
{
    typedef RLayout<NT, SYMMETRIC> LO;
    constexpr int TCAP = LO::kTileCap, ARR = LO::kArr;
    extern __shared__ __align__(16) int s_mem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* const wmem = s_mem + warp * LO::kWarpInts;
    float* const xyz = reinterpret_cast<float*>(wmem);                  // [NT][3][ARR]
    float* const tr2 = reinterpret_cast<float*>(wmem + LO::kOffR2);     // [NT][ARR]
    int* const tid = wmem + LO::kOffId;                                 // [NT][TCAP]
    const unsigned char* const lists = reinterpret_cast<const unsigned char*>(wmem + LO::kOffLists);
    int* const run_base = wmem + LO::kOffRuns;
    int* const stage = wmem;                                            // aliases xyz: written only after the tiles are dead
    const unsigned list_e0 = smem_addr_u32(wmem + LO::kOffLists) + lane;    // even candidates; row k of lane l at byte k*32 + l
    const unsigned list_o0 = list_e0 + LO::kCapSub * 32;                     // odd candidates
    const unsigned lim_e = list_e0 + 32 * (LO::kCapSub - 8), lim_o = list_o0 + 32 * (LO::kCapSub - 8);

    const unsigned lt = lanemask_lt();
    const int query_limit = a.query_limit;
    const bool same_set = a.same_set != 0;
    const uint32_t n_cells = (uint32_t)a.n_q_cells;
    const float qnan = __int_as_float(0x7fc00000);      // a lane without a query / a padding candidate: d2 = NaN, never a hit
    unsigned nb_sum = 0;

    uint32_t c0 = 0;
    if (lane == 0) c0 = atomicAdd(a.ticket, (uint32_t)kRoundCells);
    c0 = __shfl_sync(kFull, c0, 0);

    while (c0 < n_cells) {
        const int nb = (int)min((uint32_t)kRoundCells, n_cells - c0);
        Key my_key = 0;
        int my_start = 0;
        if (lane < nb) {
            my_key = a.q_cell_key[c0 + lane];
            my_start = (int)a.q_cell_start[c0 + lane];
        }
        const int batch_end = (int)a.q_cell_start[c0 + nb];
        uint32_t t_next = 0;
        if (lane == 0) t_next = atomicAdd(a.ticket, (uint32_t)kRoundCells);     // consumed at the end of the batch

        int ci = 0, qdone = 0;
        while (ci < nb) {
            // ---------------- form a round: up to NT consecutive cells, at most 32 queries; a cell is only split when it alone
            // has more than 32 queries (its tile is then staged once per chunk)
            int s_cell[NT], s_cellqb[NT], s_qcnt[NT], s_lbase[NT];
            int ns = 0, nq = 0;
            const int round_qb = __shfl_sync(kFull, my_start, ci) + qdone;
            bool closed = false;
#pragma unroll
            for (int s = 0; s < NT; s++) {
                s_cell[s] = 0; s_cellqb[s] = 0; s_qcnt[s] = 0; s_lbase[s] = 0;
                if (!closed && ci < nb) {
                    const int cb = __shfl_sync(kFull, my_start, ci);
                    const int ce_s = __shfl_sync(kFull, my_start, (ci + 1) & 31);
                    const int ce = (ci + 1 < nb) ? ce_s : batch_end;
                    const int rem = ce - cb - qdone;
                    if (rem <= 32 - nq) {
                        s_cell[s] = ci; s_cellqb[s] = cb; s_qcnt[s] = rem; s_lbase[s] = nq;
                        nq += rem; ci++; qdone = 0; ns = s + 1;
                    } else if (s == 0) {
                        s_cell[0] = ci; s_cellqb[0] = cb; s_qcnt[0] = 32; s_lbase[0] = 0;
                        nq = 32; qdone += 32; ns = 1; closed = true;
                    } else {
                        closed = true;
                    }
                }
            }

            // ---------------- the queries of the round are consecutive in the sorted array: lane l owns round_qb + l
            const bool has_q = lane < nq;
            const int qpos = round_qb + lane;
            float4 q = make_float4(qnan, 0.0f, 0.0f, __int_as_float(0x7fffffff));
            float r2 = -1.0f;
            if (has_q) {
                q = a.q_pts[qpos];
                r2 = VARIABLE ? a.q_r2[qpos] : a.r2_fixed;
            }

            // ---------------- neighbour lookups of all cells of the round in flight together
            NeighborLookup<Key, DENSE> L[NT];
#pragma unroll
            for (int s = 0; s < NT; s++)
                if (s < ns) lookup_issue(L[s], a, __shfl_sync(kFull, my_key, s_cell[s]), lane);

            // ---------------- stage the candidate tile of every cell (structure of arrays)
            __syncwarp();                       // the previous round's flush has finished reading the staging buffer
            int T[NT], self_pre[NT];
            unsigned slow_mask = 0;
            int maxT = 0;
#pragma unroll
            for (int s = 0; s < NT; s++) {
                T[s] = 0; self_pre[s] = 0;
                if (s < ns) {
                    int rs, rc, pre;
                    lookup_resolve(L[s], a, rs, rc);
                    __syncwarp();               // the previous cell's run table is no longer read
                    const int Ts = build_runs(rs, rc, run_base, lane, lt, pre);
                    T[s] = Ts;
                    self_pre[s] = __shfl_sync(kFull, pre, 13);     // lane 13 = offset (0,0,0): the cell itself when same_set
                    if (Ts > TCAP) {
                        slow_mask |= 1u << s;
                    } else {
                        maxT = max(maxT, Ts);
                        float* const X = xyz + (s * 3) * ARR;
                        float* const Y = X + ARR;
                        float* const Z = Y + ARR;
                        float* const R = tr2 + s * ARR;
                        int* const ID = tid + s * TCAP;
#pragma unroll
                        for (int h = 0; h < TCAP / 256; h++) {
                            if (h * 256 < Ts) {
                                int pos[8];
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const int sb = h * 256 + j * 32;
                                    pos[j] = 0;
                                    if (sb < Ts) pos[j] = candidate_pos(pre, rc, run_base, sb, lane);
                                }
                                float4 v[8];
                                float w[8];
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const int t = h * 256 + j * 32 + lane;
                                    v[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                    w[j] = 0.0f;
                                    if (t < Ts) {
                                        v[j] = __ldg(a.c_pts + pos[j]);
                                        if (SYMMETRIC) w[j] = __ldg(a.c_r2 + pos[j]);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const int t = h * 256 + j * 32 + lane;
                                    if (t < Ts) {
                                        X[t] = v[j].x; Y[t] = v[j].y; Z[t] = v[j].z;
                                        ID[t] = __float_as_int(v[j].w);
                                        if (SYMMETRIC) R[t] = w[j];
                                    }
                                }
                            }
                        }
                    }
                }
            }
            // pad every tile with never-hit candidates up to the common (even, unrolled) trip count of the round
            const int Tpad = (maxT + 7) & ~7;
#pragma unroll
            for (int s = 0; s < NT; s++) {
                if (s < ns && !((slow_mask >> s) & 1u)) {
                    float* const X = xyz + (s * 3) * ARR;
                    for (int t = T[s] + lane; t < Tpad; t += 32) {
                        X[t] = qnan; X[ARR + t] = 0.0f; X[2 * ARR + t] = 0.0f;
                        if (SYMMETRIC) tr2[s * ARR + t] = -1.0f;
                    }
                }
            }

            // ---------------- which tile is mine
            int my_s = 0, my_selfpre = 0, my_cellqb = 0;
#pragma unroll
            for (int s = 0; s < NT; s++) {
                if (lane >= s_lbase[s] && lane < s_lbase[s] + s_qcnt[s]) { my_s = s; my_selfpre = self_pre[s]; my_cellqb = s_cellqb[s]; }
            }
            const int qidx = __float_as_int(q.w);
            bool active = has_q && qidx < query_limit && !((slow_mask >> my_s) & 1u);
            if (!active) { q.x = qnan; r2 = -1.0f; }
            __syncwarp();                       // tiles complete

            // ---------------- inner loop: two candidates per step, private hit lists, no cross-lane traffic
            const unsigned long long* const X2 = reinterpret_cast<const unsigned long long*>(xyz + (my_s * 3) * ARR);
            const unsigned long long* const Y2 = X2 + ARR / 2;
            const unsigned long long* const Z2 = Y2 + ARR / 2;
            const unsigned long long* const R2 = reinterpret_cast<const unsigned long long*>(tr2 + my_s * ARR);
            const f32x2 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
            const int npairs = Tpad >> 1;       // multiple of 4
            unsigned pe = list_e0, po = list_o0;
            bool ovf = false;
            struct Quad { f32x2 x[4], y[4], z[4], r[4]; };
            auto load_quad = [&](Quad& c, int it) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    c.x[u] = X2[it + u]; c.y[u] = Y2[it + u]; c.z[u] = Z2[it + u];
                    c.r[u] = SYMMETRIC ? R2[it + u] : 0ull;
                }
            };
            // four pairs from registers; the NEXT four are loaded first (reads past Tpad stay inside the warp's shared memory
            // block and are never used)
            auto step = [&](const Quad& c, Quad& nxt, int it) {
                load_quad(nxt, it + 4);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const f32x2 dx = sub2(qx, c.x[u]);
                    const f32x2 dy = sub2(qy, c.y[u]);
                    const f32x2 dz = sub2(qz, c.z[u]);
                    const f32x2 d2p = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
                    float d2l, d2h;
                    unpack2(d2p, d2l, d2h);
                    bool hl = d2l <= r2, hh = d2h <= r2;
                    if (SYMMETRIC) {
                        float rl, rh;
                        unpack2(c.r[u], rl, rh);
                        hl = hl || (d2l <= rl);
                        hh = hh || (d2h <= rh);
                    }
                    if (hl) { sts_u8(pe, it + u); pe += 32; }
                    if (hh) { sts_u8(po, it + u); po += 32; }
                }
            };
            Quad quad_a, quad_b;
            load_quad(quad_a, 0);
            for (int it = 0; it < npairs; it += 8) {
                step(quad_a, quad_b, it);
                if (it + 4 < npairs) step(quad_b, quad_a, it + 4);
                // a private list is about to run out of rows (8 more could be needed per trip): redo the round on the slow path
                if (__any_sync(kFull, pe > lim_e || po > lim_o)) { ovf = true; break; }
            }
            __syncwarp();                       // every lane is done with the tiles; the private lists are complete

            // ---------------- expand the private lists into [n, j0, j1, ...] and flush
            if (ovf) { slow_mask = (1u << ns) - 1u; active = false; }
            const int ne = (int)(pe - list_e0) >> 5, no = (int)(po - list_o0) >> 5;
            const int n = active ? ne + no - (same_set ? 1 : 0) : 0;
            const int len = active ? n + 1 : 0;
            const int inc = warp_inclusive_scan(len, lane);
            const int total = __shfl_sync(kFull, inc, 31);
            const int off = inc - len;
            if (total > 0) {
                const int w4 = (total + 3) & ~3;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.cursor, (unsigned long long)w4);
                const int tself = same_set ? my_selfpre + (qpos - my_cellqb) : -1;
                const int* const ID = tid + my_s * TCAP;
                auto write_list = [&](int* dst) {
                    if (active) {
                        dst[0] = n;
                        int w = 1;
                        for (int k = 0; k < ne; k++) {
                            const int t = 2 * (int)lists[k * 32 + lane];
                            if (t != tself) dst[w++] = ID[t];
                        }
                        for (int k = 0; k < no; k++) {
                            const int t = 2 * (int)lists[LO::kCapSub * 32 + k * 32 + lane] + 1;
                            if (t != tself) dst[w++] = ID[t];
                        }
                    }
                };
                const bool staged = total <= LO::kStageInts;
                if (staged) write_list(stage + off);
                base = __shfl_sync(kFull, base, 0);
                if ((long long)(base + w4) <= a.capacity) {
                    if (staged) {
                        __syncwarp();
                        const int4* src = reinterpret_cast<const int4*>(stage);
                        int4* dst4 = reinterpret_cast<int4*>(a.ragged + base);
                        for (int t = lane; t < (w4 >> 2); t += 32) st_stream_i4(dst4 + t, src[t]);
                    } else {
                        write_list(a.ragged + base + off);     // more ids than the staging buffer holds (very rare): straight to HBM
                    }
                    if (active) a.list_pos[qidx] = (long long)base + off;
                    nb_sum += (unsigned)__reduce_add_sync(kFull, n);
                } else if (lane == 0) {
                    *a.overflow = 1;
                }
            }

            // ---------------- cells that did not fit a tile / rounds that overflowed a private list
            while (slow_mask) {
                const int s = __ffs((int)slow_mask) - 1;
                slow_mask &= slow_mask - 1u;
                const Key ck = __shfl_sync(kFull, my_key, pick(s_cell, s));
                const int qb = round_qb + pick(s_lbase, s);
                nb_sum += slow_queries<Key, VARIABLE, SYMMETRIC, DENSE>(a, ck, qb, qb + pick(s_qcnt, s), pick(s_cellqb, s), run_base, lane, lt);
            }
        }
        if (nb_sum > 0x40000000u) {
            if (lane == 0) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
            nb_sum = 0;
        }
        c0 = __shfl_sync(kFull, t_next, 0);
    }
    if (lane == 0 && nb_sum) atomicAdd(a.n_neighbors, (unsigned long long)nb_sum);
}

}  // namespace tnsb
