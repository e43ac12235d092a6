// query_rounds.cuh -- second generation of the 27-cell fixed-radius query ("round" kernel).  Same contract and arguments as
// query_kernel (query.cuh); replaces _solve_leaves / _prepare_brute_force[_simd] / _brute_force[_simd] of the reference
// (TreeNSearch.cpp:1823-1872, :2161-2399, :2400-2569).
//
// Why a second kernel: query_kernel keeps the candidates in registers and gives every LANE a candidate, so each
// (query, 32 candidates) step needs a ballot, two popcounts and an address computation to compact the hits -- ncu shows it
// bound by instruction issue (~190 warp instructions per query at 10M points), not by memory.  Here the roles are swapped:
//
//   * every LANE owns a QUERY.  A warp forms a "round" of up to 32 consecutive query points taken from up to NT consecutive
//     occupied cells of the sorted grid;
//   * the grid is sorted by ROW KEYS (common.cuh RowKey: x consecutive inside a (y, z) row), so the 27-cell neighbourhood of a
//     cell is 9 contiguous runs of the sorted point array, and the run bounds are 18 loads from a prefix table that answers
//     "first point of cell k" for every cell, empty or not (a hash of the occupied cells for huge sparse domains);
//   * the 9 runs of each cell of the round are staged ONCE into a shared-memory tile in structure-of-arrays form X[t], Y[t],
//     Z[t], ID[t] (coalesced 16-byte loads, conflict free 4-byte stores); the lookups of the NEXT round are already in flight;
//   * inner loop: a lane walks the tile of ITS cell -- three LDS.128 bring four candidates as packed pairs (x0,x1)(x2,x3) ...,
//     six packed FADD2/FMUL2/FFMA2 per pair give both distances in the reference's exact arithmetic
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
    static constexpr int kArr = kTileCap + 4;             // words per coordinate array; the 4 extra words stagger the banks of the NT tiles
    static constexpr int kStageInts = NT * 3 * kArr;      // X/Y/Z of all tiles; doubles as the output staging buffer
    static constexpr int kOffR2 = kStageInts;             // candidate r^2 (symmetric variable radius only)
    static constexpr int kOffId = kOffR2 + (SYMMETRIC ? NT * kArr : 0);
    static constexpr int kListStride = 84;                // bytes per lane and sub-list (21 words: odd, so 32 lanes hit 32 banks)
    static constexpr int kOffLists = kOffId + NT * kTileCap;
    static constexpr int kOffRuns = kOffLists + 2 * 32 * kListStride / 4;    // int4 {start, count, first candidate number, -} x 9 x NT
    static constexpr int kWarpInts = kOffRuns + NT * 9 * 4;
    static constexpr int kWarps = SYMMETRIC ? 8 : 10;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kBytes = kWarps * kWarpInts * 4;
    static_assert(kArr % 4 == 0 && kOffR2 % 4 == 0 && kOffRuns % 4 == 0 && kWarpInts % 4 == 0, "16-byte alignment of the vector accesses");
    static_assert(kTileCap / 2 <= 256, "a pair number must fit one byte");
    static_assert(kBytes <= 227 * 1024, "shared memory per SM");
};

// pending neighbour lookup of ONE cell.  Table mode: lanes 0..8 hold the first point of the 9 row runs, lanes 9..17 their ends.
// Hash mode: lanes 0..26 hold one neighbour cell each (3 per row).
template <typename Key, bool DENSE>
struct RowLookup {
    uint32_t v;
    Key key;
    uint32_t slot;
    bool valid;
    typename HashSlot<Key>::Raw e;
};

template <typename Key, bool DENSE>
__device__ __forceinline__ void row_lookup_issue(RowLookup<Key, DENSE>& L, const QueryArgs<Key>& a, Key cell_key, int lane)
{
    const int bits = a.bits;
    const Key xmask = (Key)(((Key)1 << bits) - 1);
    const int x = (int)(cell_key & xmask);
    const Key row = cell_key >> bits;
    const Key row_mask = a.key_mask >> bits;
    int l = lane;
    asm volatile("" : "+r"(l));       // keeps the per-lane neighbour constants out of long-lived registers
    L.v = 0;
    L.key = 0;
    L.slot = 0;
    if (DENSE) {
        const int r = l >= 9 ? l - 9 : l;
        bool valid = l < 18;
        const Key nrow = row_neighbor<Key>(row, r % 3 - 1, r / 3 - 1, row_mask, valid);
        const int xl = max(x - 1, 0), xh = min(x + 1, (int)xmask) + 1;
        const Key idx = (nrow << bits) + (Key)(l >= 9 ? xh : xl);
        L.valid = valid;
        if (valid) L.v = __ldg(a.first + idx);
    } else {
        const int r = l / 3, xx = x + (l % 3) - 1;
        bool valid = l < 27;
        const Key nrow = row_neighbor<Key>(row, r % 3 - 1, r / 3 - 1, row_mask, valid);
        valid = valid && xx >= 0 && xx <= (int)xmask;
        L.valid = valid;
        L.key = (nrow << bits) | (Key)(valid ? xx : 0);
        L.slot = Morton<Key>::hash(L.key) >> (32 - a.hash_log2);
        if (valid) L.e = HashSlot<Key>::load(a.htable, L.slot);
    }
}

// after this, lanes 0..8 hold (start, count) of the 9 row runs of the cell; the other lanes hold (0, 0)
template <typename Key, bool DENSE>
__device__ __forceinline__ void row_lookup_resolve(const RowLookup<Key, DENSE>& L, const QueryArgs<Key>& a, int lane, int& rs, int& rc)
{
    if (DENSE) {
        const uint32_t hi = __shfl_down_sync(kFull, L.v, 9);
        rs = lane < 9 ? (int)L.v : 0;
        rc = lane < 9 ? (int)(hi - L.v) : 0;
    } else {
        int s = 0, c = 0;
        if (L.valid) {
            const uint32_t hmask = (1u << a.hash_log2) - 1u;
            typename HashSlot<Key>::Raw e = L.e;
            uint32_t slot = L.slot;
            for (;;) {
                if (HashSlot<Key>::matches(e, L.key)) { s = HashSlot<Key>::start(e); c = HashSlot<Key>::count(e); break; }
                if (HashSlot<Key>::is_empty(e)) break;
                slot = (slot + 1) & hmask;
                e = HashSlot<Key>::load(a.htable, slot);
            }
        }
        // the three cells of a row are consecutive keys: their points are contiguous, the run starts at the first non-empty one
        const int c1 = __shfl_down_sync(kFull, c, 1), c2 = __shfl_down_sync(kFull, c, 2);
        const int s1 = __shfl_down_sync(kFull, s, 1), s2 = __shfl_down_sync(kFull, s, 2);
        const int cnt3 = c + c1 + c2;
        const int st3 = c > 0 ? s : (c1 > 0 ? s1 : s2);
        const int src = (3 * lane) & 31;
        const int rs9 = __shfl_sync(kFull, st3, src), rc9 = __shfl_sync(kFull, cnt3, src);
        rs = lane < 9 ? rs9 : 0;
        rc = lane < 9 ? rc9 : 0;
    }
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
// list): two sweeps per query over the 9 runs through L1/L2, lists written straight to the ragged buffer.
template <typename Key, bool VARIABLE, bool SYMMETRIC, bool DENSE>
__device__ __forceinline__ unsigned slow_queries(const QueryArgs<Key>& a, Key cell_key, int qb, int qe, int lane, unsigned lt)
{
    RowLookup<Key, DENSE> L;
    row_lookup_issue(L, a, cell_key, lane);
    int rs, rc;
    row_lookup_resolve(L, a, lane, rs, rc);
    unsigned found = 0;
    for (int qi = qb; qi < qe; qi++) {
        const float4 qv = a.q_pts[qi];
        const int qidx = __float_as_int(qv.w);
        if (qidx >= a.query_limit) continue;
        const float r2 = VARIABLE ? a.q_r2[qi] : a.r2_fixed;
        const int self_pos = a.same_set ? qi : -1;         // same set: query and candidate arrays are the same sorted array
        int n = 0;
        for (int r = 0; r < 9; r++) {
            const int s0 = __shfl_sync(kFull, rs, r), cnt = __shfl_sync(kFull, rc, r);
            for (int t0 = 0; t0 < cnt; t0 += 32) {
                const int pos = s0 + t0 + lane;
                bool h = false;
                if (t0 + lane < cnt && pos != self_pos) {
                    const float4 v = a.c_pts[pos];
                    const float d2 = dist2(qv.x, qv.y, qv.z, v.x, v.y, v.z);
                    h = d2 <= r2;
                    if (SYMMETRIC) h = h || (d2 <= a.c_r2[pos]);
                }
                n += __popc(__ballot_sync(kFull, h));
            }
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
        for (int r = 0; r < 9; r++) {
            const int s0 = __shfl_sync(kFull, rs, r), cnt = __shfl_sync(kFull, rc, r);
            for (int t0 = 0; t0 < cnt; t0 += 32) {
                const int pos = s0 + t0 + lane;
                bool h = false;
                int id = -1;
                if (t0 + lane < cnt && pos != self_pos) {
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
        }
        found += (unsigned)n;
    }
    return found;
}

template <int NT>
struct RoundPlan {
    int ns, nq, round_qb;
    int s_cell[NT], s_cellqb[NT], s_qcnt[NT], s_lbase[NT];
};

template <typename Key, int NT, bool VARIABLE, bool SYMMETRIC, bool DENSE>
__global__ void __launch_bounds__(RLayout<NT, SYMMETRIC>::kThreads, 1) query_rounds_kernel(const QueryArgs<Key> a)
{
    typedef RLayout<NT, SYMMETRIC> LO;
    constexpr int TCAP = LO::kTileCap, ARR = LO::kArr, LSTRIDE = LO::kListStride;
    extern __shared__ __align__(16) int s_mem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* const wmem = s_mem + warp * LO::kWarpInts;
    float* const xyz = reinterpret_cast<float*>(wmem);                  // [NT][3][ARR]
    float* const tr2 = reinterpret_cast<float*>(wmem + LO::kOffR2);     // [NT][ARR]
    int* const tid = wmem + LO::kOffId;                                 // [NT][TCAP]
    const uint32_t* const list_e = reinterpret_cast<const uint32_t*>(wmem + LO::kOffLists) + lane * (LSTRIDE / 4);      // even candidates
    const uint32_t* const list_o = list_e + 32 * (LSTRIDE / 4);                                                         // odd candidates
    int4* const runtab = reinterpret_cast<int4*>(wmem + LO::kOffRuns);  // [NT][9]
    int* const stage = wmem;                                            // aliases xyz: written only after the tiles are dead
    const unsigned list_e0 = smem_addr_u32(list_e), list_o0 = smem_addr_u32(list_o);
    const unsigned lim_e = list_e0 + (LSTRIDE - 8), lim_o = list_o0 + (LSTRIDE - 8);

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
        // a round: up to NT consecutive cells, at most 32 queries; a cell is only split when it alone has more than 32 queries
        // (its tile is then staged once per chunk).  ns == 0: the batch is exhausted.
        auto form_round = [&](RoundPlan<NT>& p) {
            p.ns = 0; p.nq = 0;
            p.round_qb = __shfl_sync(kFull, my_start, ci & 31) + qdone;
            bool closed = false;
#pragma unroll
            for (int s = 0; s < NT; s++) {
                p.s_cell[s] = 0; p.s_cellqb[s] = 0; p.s_qcnt[s] = 0; p.s_lbase[s] = 0;
                if (!closed && ci < nb) {
                    const int cb = __shfl_sync(kFull, my_start, ci);
                    const int ce_s = __shfl_sync(kFull, my_start, (ci + 1) & 31);
                    const int ce = (ci + 1 < nb) ? ce_s : batch_end;
                    const int rem = ce - cb - qdone;
                    if (rem <= 32 - p.nq) {
                        p.s_cell[s] = ci; p.s_cellqb[s] = cb; p.s_qcnt[s] = rem; p.s_lbase[s] = p.nq;
                        p.nq += rem; ci++; qdone = 0; p.ns = s + 1;
                    } else if (s == 0) {
                        p.s_cell[0] = ci; p.s_cellqb[0] = cb; p.s_qcnt[0] = 32; p.s_lbase[0] = 0;
                        p.nq = 32; qdone += 32; p.ns = 1; closed = true;
                    } else {
                        closed = true;
                    }
                }
            }
        };
        // neighbour lookups of all cells of a round + its query points: issued one round ahead
        auto issue_round = [&](const RoundPlan<NT>& p, RowLookup<Key, DENSE> (&L)[NT], float4& q, float& r2) {
#pragma unroll
            for (int s = 0; s < NT; s++)
                if (s < p.ns) row_lookup_issue(L[s], a, __shfl_sync(kFull, my_key, p.s_cell[s]), lane);
            q = make_float4(qnan, 0.0f, 0.0f, __int_as_float(0x7fffffff));
            r2 = -1.0f;
            if (lane < p.nq) {
                q = a.q_pts[p.round_qb + lane];
                r2 = VARIABLE ? a.q_r2[p.round_qb + lane] : a.r2_fixed;
            }
        };

        RoundPlan<NT> plan;
        RowLookup<Key, DENSE> L[NT];
        float4 q;
        float r2;
        form_round(plan);
        issue_round(plan, L, q, r2);

        while (plan.ns > 0) {
            const int ns = plan.ns;
            // ---------------- run table of every cell: 9 x {start, count, first candidate number}
            __syncwarp();                       // the previous round's flush has finished reading the staging buffer / run table
            int T[NT], selfbase[NT];
            unsigned slow_mask = 0, long_mask = 0;
            int maxT = 0;
#pragma unroll
            for (int s = 0; s < NT; s++) {
                T[s] = 0; selfbase[s] = 0;
                if (s < ns) {
                    int rs, rc;
                    row_lookup_resolve(L[s], a, lane, rs, rc);
                    const int inc = warp_inclusive_scan(rc, lane);
                    const int Ts = __shfl_sync(kFull, inc, 31);
                    const int pre = inc - rc;
                    T[s] = Ts;
                    selfbase[s] = __shfl_sync(kFull, pre - rs, 4);      // run 4 = the cell's own row: candidate number of sorted position p is selfbase + p
                    if (lane < 9) runtab[s * 9 + lane] = make_int4(rs, rc, pre, 0);
                    if (__any_sync(kFull, rc > 32)) long_mask |= 1u << s;
                    if (Ts > TCAP) slow_mask |= 1u << s;
                    else maxT = max(maxT, Ts);
                }
            }
            __syncwarp();
            const int Tpad = (maxT + 7) & ~7;   // common trip count of the round: a multiple of 8 candidates

            // ---------------- stage the candidate tiles (structure of arrays), two cells' loads in flight together
#pragma unroll
            for (int s0 = 0; s0 < NT; s0 += 2) {
                float4 v[2][9];
                float w[2][9];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int s = s0 + h;
                    if (s < NT && s < ns && !((slow_mask >> s) & 1u)) {
#pragma unroll
                        for (int r = 0; r < 9; r++) {
                            const int4 t = runtab[s * 9 + r];
                            if (lane < t.y) {
                                v[h][r] = __ldg(a.c_pts + t.x + lane);
                                if (SYMMETRIC) w[h][r] = __ldg(a.c_r2 + t.x + lane);
                            }
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int s = s0 + h;
                    if (s < NT && s < ns && !((slow_mask >> s) & 1u)) {
                        float* const X = xyz + (s * 3) * ARR;
                        float* const Y = X + ARR;
                        float* const Z = Y + ARR;
                        float* const R = tr2 + s * ARR;
                        int* const ID = tid + s * TCAP;
#pragma unroll
                        for (int r = 0; r < 9; r++) {
                            const int4 t = runtab[s * 9 + r];
                            if (lane < t.y) {
                                const int k = t.z + lane;
                                X[k] = v[h][r].x; Y[k] = v[h][r].y; Z[k] = v[h][r].z;
                                ID[k] = __float_as_int(v[h][r].w);
                                if (SYMMETRIC) R[k] = w[h][r];
                            }
                        }
                        if ((long_mask >> s) & 1u) {
                            // runs longer than one warp (dense rows): the remaining elements
                            for (int r = 0; r < 9; r++) {
                                const int4 t = runtab[s * 9 + r];
                                for (int e = 32 + lane; e < t.y; e += 32) {
                                    const float4 u = __ldg(a.c_pts + t.x + e);
                                    const int k = t.z + e;
                                    X[k] = u.x; Y[k] = u.y; Z[k] = u.z;
                                    ID[k] = __float_as_int(u.w);
                                    if (SYMMETRIC) R[k] = __ldg(a.c_r2 + t.x + e);
                                }
                            }
                        }
                        // pad with never-hit candidates (x = NaN is enough) up to the common trip count of the round
                        for (int k = T[s] + lane; k < Tpad; k += 32) X[k] = qnan;
                    }
                }
            }

            // ---------------- which tile is mine
            const int nq = plan.nq, round_qb = plan.round_qb;
            int my_s = 0, my_selfbase = 0;
#pragma unroll
            for (int s = 0; s < NT; s++) {
                if (lane >= plan.s_lbase[s] && lane < plan.s_lbase[s] + plan.s_qcnt[s]) { my_s = s; my_selfbase = selfbase[s]; }
            }
            const int qpos = round_qb + lane;
            const int qidx = __float_as_int(q.w);
            bool active = lane < nq && qidx < query_limit && !((slow_mask >> my_s) & 1u);
            const float qx1 = active ? q.x : qnan, r2q = active ? r2 : -1.0f;
            const f32x2 qx = pack2(qx1, qx1), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);

            // ---------------- next round: plan, lookups and query points go in flight now and land during the inner loop
            const RoundPlan<NT> cur = plan;
            form_round(plan);
            issue_round(plan, L, q, r2);
            __syncwarp();                       // tiles complete

            // ---------------- inner loop: private hit lists, no cross-lane traffic
            const ulonglong2* const X4 = reinterpret_cast<const ulonglong2*>(xyz + (my_s * 3) * ARR);
            const ulonglong2* const Y4 = X4 + ARR / 4;
            const ulonglong2* const Z4 = Y4 + ARR / 4;
            const ulonglong2* const R4 = reinterpret_cast<const ulonglong2*>(tr2 + my_s * ARR);
            const int npairs = Tpad >> 1;       // multiple of 4
            unsigned pe = list_e0, po = list_o0;
            bool ovf = false;
            struct Quad { f32x2 x[4], y[4], z[4], r[4]; };
            // four pairs = eight candidates: two LDS.128 per coordinate
            auto load_quad = [&](Quad& c, int it) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const ulonglong2 vx = X4[(it >> 1) + u], vy = Y4[(it >> 1) + u], vz = Z4[(it >> 1) + u];
                    c.x[2 * u] = vx.x; c.x[2 * u + 1] = vx.y;
                    c.y[2 * u] = vy.x; c.y[2 * u + 1] = vy.y;
                    c.z[2 * u] = vz.x; c.z[2 * u + 1] = vz.y;
                    if (SYMMETRIC) {
                        const ulonglong2 vr = R4[(it >> 1) + u];
                        c.r[2 * u] = vr.x; c.r[2 * u + 1] = vr.y;
                    } else {
                        c.r[2 * u] = 0ull; c.r[2 * u + 1] = 0ull;
                    }
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
                    bool hl = d2l <= r2q, hh = d2h <= r2q;
                    if (SYMMETRIC) {
                        float rl, rh;
                        unpack2(c.r[u], rl, rh);
                        hl = hl || (d2l <= rl);
                        hh = hh || (d2h <= rh);
                    }
                    if (hl) { sts_u8(pe, it + u); pe += 1; }
                    if (hh) { sts_u8(po, it + u); po += 1; }
                }
            };
            Quad quad_a, quad_b;
            load_quad(quad_a, 0);
            for (int it = 0; it < npairs; it += 8) {
                step(quad_a, quad_b, it);
                if (it + 4 < npairs) step(quad_b, quad_a, it + 4);
                // a private list is about to run out of room (8 more entries could be needed per trip): redo the round on the slow path
                if (__any_sync(kFull, pe > lim_e || po > lim_o)) { ovf = true; break; }
            }
            __syncwarp();                       // every lane is done with the tiles; the private lists are complete

            // ---------------- expand the private lists into [n, j0, j1, ...] and flush
            if (ovf) { slow_mask = (1u << ns) - 1u; active = false; }
            const int ne = (int)(pe - list_e0), no = (int)(po - list_o0);
            const int n = active ? ne + no - (same_set ? 1 : 0) : 0;
            const int len = active ? n + 1 : 0;
            const int inc = warp_inclusive_scan(len, lane);
            const int total = __shfl_sync(kFull, inc, 31);
            const int off = inc - len;
            if (total > 0) {
                const int w4 = (total + 3) & ~3;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.cursor, (unsigned long long)w4);
                const int tself = same_set ? my_selfbase + qpos : -1;       // candidate number of the query itself
                const int bself_e = (tself >= 0 && !(tself & 1)) ? (tself >> 1) : -1;
                const int bself_o = (tself >= 0 && (tself & 1)) ? (tself >> 1) : -1;
                const int* const ID = tid + my_s * TCAP;
                auto write_list = [&](int* dst) {
                    if (active) {
                        dst[0] = n;
                        int w = 1;
                        for (int k = 0; k < ne; k += 4) {
                            const uint32_t pk = list_e[k >> 2];
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const int b = (int)((pk >> (8 * i)) & 0xffu);
                                if (k + i < ne && b != bself_e) dst[w++] = ID[2 * b];
                            }
                        }
                        for (int k = 0; k < no; k += 4) {
                            const uint32_t pk = list_o[k >> 2];
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const int b = (int)((pk >> (8 * i)) & 0xffu);
                                if (k + i < no && b != bself_o) dst[w++] = ID[2 * b + 1];
                            }
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
                const Key ck = __shfl_sync(kFull, my_key, pick(cur.s_cell, s));
                const int qb = round_qb + pick(cur.s_lbase, s);
                nb_sum += slow_queries<Key, VARIABLE, SYMMETRIC, DENSE>(a, ck, qb, qb + pick(cur.s_qcnt, s), lane, lt);
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
