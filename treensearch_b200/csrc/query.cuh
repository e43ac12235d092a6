// query.cuh -- the 27-cell fixed-radius distance query.  Replaces _solve_leaves / _prepare_brute_force[_simd] /
// _brute_force[_simd] of the reference (TreeNSearch.cpp:1823-1872, :2161-2399, :2400-2569).
//
// Work decomposition (one launch per active ordered pair set_i -> set_j):
//   * a task is one occupied cell of set_i; a warp pulls batches of consecutive (Morton ordered) cells from a ticket counter;
//   * lanes 0..26 look the 27 neighbour cells of set_j up in the cell hash and a warp scan turns their populations into a
//     dense candidate list; every lane then owns candidate t = slot*32 + lane and keeps it in REGISTERS for the whole cell
//     (one coalesced 16-byte load per candidate per cell, instead of one per (query, candidate) pair);
//   * for each query point of the cell (broadcast from shared memory) every lane tests its candidates, two per instruction
//     with Blackwell's packed FADD2/FMUL2/FFMA2, in the reference's exact arithmetic
//     d2 = fma(dz,dz, fma(dx,dx, dy*dy)) <= r^2  (TreeNSearch.cpp:2477-2486 as compiled); a warp ballot per
//     slot compacts the hits straight into a per-warp shared-memory staging buffer as  [n, j0, j1, ...]  (the reference's
//     list layout, TreeNSearch.h:395) -- no shuffles and no dependent scan chain in the inner loop;
//   * when the staging buffer is full the warp reserves a range of the global ragged buffer with ONE atomicAdd, copies the
//     staged lists with fully coalesced stores and publishes list_pos[i] for the staged queries.
// Lists are therefore written exactly once, in one pass (no count pass), and nothing is ever re-read from HBM.
//
// Self exclusion: only the identical (set, index) is excluded (TreeNSearch.cpp:2464-2466); coincident points are neighbours.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace tnsb {

constexpr int kQueryThreads = 256;
constexpr int kQueryWarps = kQueryThreads / 32;
#ifndef TNSB_QUERY_BLOCKS_PER_SM
#define TNSB_QUERY_BLOCKS_PER_SM 3
#endif
constexpr int kQueryBlocksPerSM = TNSB_QUERY_BLOCKS_PER_SM;
constexpr int kStageInts = 1536;          // per-warp staging capacity (ints)
constexpr int kStageRecs = 192;           // per-warp staged list records
constexpr int kCellsPerTicket = 16;
// per-warp shared memory (ints): stage | rec_idx | rec_off | run_base[32] | (32 spare) | query float4[32] | query r2[32]
constexpr int kOffRecIdx = kStageInts;
constexpr int kOffRecOff = kOffRecIdx + kStageRecs;
constexpr int kOffRunStart = kOffRecOff + kStageRecs;
constexpr int kOffRunPre = kOffRunStart + 32;
constexpr int kOffQbuf = kOffRunPre + 32;        // must be a multiple of 4 ints (float4 alignment)
constexpr int kOffQr2 = kOffQbuf + 128;
constexpr int kWarpSmemInts = kOffQr2 + 32;
constexpr int kQuerySmemBytes = kQueryWarps * kWarpSmemInts * 4;
static_assert(kOffQbuf % 4 == 0 && kWarpSmemInts % 4 == 0, "float4 alignment of the per-warp query buffer");

template <typename Key>
struct QueryArgs {
    // searching set (set_i)
    const float4* q_pts;          // sorted (x, y, z, bits(index))
    const float* q_r2;            // sorted r^2 (variable radius mode)
    const Key* q_cell_key;
    const uint32_t* q_cell_start; // n_q_cells + 1
    int n_q_cells;
    int query_limit;              // points with index >= limit are find-only (INT_MAX: none)
    // searched set (set_j)
    const float4* c_pts;
    const float* c_r2;
    const typename HashSlot<Key>::Raw* htable;   // cell key -> [start, end) of the cell's run in c_pts (sparse / huge domains)
    const uint2* dense;           // Morton-indexed direct table {start, end} (domains of <= 2^27 cells): no probing at all
    int hash_log2;
    int same_set;
    Key key_mask;                 // (1 << 3*bits) - 1
    float r2_fixed;
    // output
    int32_t* ragged;
    long long capacity;
    long long* list_pos;
    unsigned long long* cursor;   // next free int of the ragged buffer (always a multiple of 4)
    uint32_t* ticket;
    unsigned long long* n_neighbors;
    int* overflow;
};

struct WarpStage {
    int* ints;       // [kStageInts]  staged lists  [n, j0, j1, ...]
    int* rec_idx;    // [kStageRecs]  query index of every staged list
    int* rec_off;    // [kStageRecs]  its offset inside ints
    int* run_base;   // [32]          (start - prefix) of the non-empty neighbour runs, compacted
    int wpos;
    int nrec;
    unsigned nb_sum;              // neighbour ids written by this warp (flushed to the 64-bit global counter before it can wrap)
};

// write-once data: streaming stores keep the lists from evicting the candidate tiles out of L2
__device__ __forceinline__ void st_stream_i4(int4* p, const int4& v) { __stcs(p, v); }

// Flush the staged lists: ONE atomicAdd reserves a 16-byte aligned range of the ragged buffer, the copy runs as 128-bit
// loads / stores, then list_pos is published for the staged queries.
template <typename Key>
__device__ __forceinline__ void stage_flush(WarpStage& st, const QueryArgs<Key>& a, int lane)
{
    __syncwarp();
    if (st.wpos > 0) {
        const int w4 = (st.wpos + 3) & ~3;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.cursor, (unsigned long long)w4);
        base = __shfl_sync(kFull, base, 0);
        if ((long long)(base + w4) <= a.capacity) {
            const int4* src = reinterpret_cast<const int4*>(st.ints);
            int4* dst = reinterpret_cast<int4*>(a.ragged + base);
            for (int t = lane; t < (w4 >> 2); t += 32) st_stream_i4(dst + t, src[t]);
            for (int k = lane; k < st.nrec; k += 32) a.list_pos[st.rec_idx[k]] = (long long)base + st.rec_off[k];
        } else if (lane == 0) {
            *a.overflow = 1;
        }
        st.nb_sum += (unsigned)(st.wpos - st.nrec);
    }
    st.wpos = 0;
    st.nrec = 0;
    __syncwarp();
}

// general path only: reserve room for one list of n ids whose size is already known.  Returns the destination of the count
// word: inside the staging buffer, or -- for lists that do not fit the staging buffer at all -- directly in the ragged buffer.
template <typename Key>
__device__ __forceinline__ int* reserve_list(WarpStage& st, const QueryArgs<Key>& a, int lane, int qidx, int n, bool& ok)
{
    ok = true;
    if (n + 1 > kStageInts) {
        const unsigned long long need = (unsigned long long)((n + 1 + 3) & ~3);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.cursor, need);
        base = __shfl_sync(kFull, base, 0);
        if ((long long)(base + need) > a.capacity) {
            if (lane == 0) *a.overflow = 1;
            ok = false;
            return nullptr;
        }
        if (lane == 0) {
            a.ragged[base] = n;
            a.list_pos[qidx] = (long long)base;
        }
        st.nb_sum += (unsigned)n;
        if (st.nb_sum > 0x40000000u) { if (lane == 0) atomicAdd(a.n_neighbors, (unsigned long long)st.nb_sum); st.nb_sum = 0; }
        return a.ragged + base;
    }
    if (st.wpos + n + 1 > kStageInts || st.nrec == kStageRecs) stage_flush(st, a, lane);
    int* dst = st.ints + st.wpos;
    if (lane == 0) {
        dst[0] = n;
        st.rec_idx[st.nrec] = qidx;
        st.rec_off[st.nrec] = st.wpos;
    }
    st.wpos += n + 1;
    st.nrec += 1;
    return dst;
}

__device__ __forceinline__ int warp_inclusive_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// the reference's distance, with explicit roundings (the library is also built with -fmad=false)
__device__ __forceinline__ float dist2(float qx, float qy, float qz, float cx, float cy, float cz)
{
    const float dx = __fsub_rn(qx, cx);
    const float dy = __fsub_rn(qy, cy);
    const float dz = __fsub_rn(qz, cz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Blackwell packed fp32 pairs (FADD2 / FMUL2 / FFMA2): two candidates per instruction, each half rounded to nearest exactly
// like the scalar instruction, so the result is bit-identical to dist2().
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 v;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// x + (-0.0) == x for every x: an exact identity that ptxas cannot fold away, used once per cell to make the candidate
// pairs live in aligned 64-bit registers (otherwise the halves stay in the LDG.128 destination registers and every FADD2 of
// the inner loop needs two extra moves to assemble its operand).
__device__ __forceinline__ f32x2 settle2(f32x2 a)
{
    f32x2 r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(0x8000000080000000ull));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// The dense candidate list of a cell is the concatenation of its (up to 27) non-empty neighbour runs.  RunTable answers
// "which position of the sorted array is candidate t" for the 32 consecutive candidates of one slot with one REDUX.OR,
// one VOTE and a popc per lane (no per-lane search): bit b of `starts` says that a run begins at candidate base+b.
struct RunTable {
    int pre;         // this lane's run: first candidate number (exclusive prefix of the run lengths)
    int cnt;         // this lane's run: length (0 = no run on this lane)
    const int* run_base;
    __device__ __forceinline__ int pos(int slot_base, int lane) const
    {
        const bool starts_here = cnt > 0 && pre >= slot_base && pre < slot_base + 32;
        const unsigned starts = __reduce_or_sync(kFull, starts_here ? (1u << (pre - slot_base)) : 0u);
        const int before = __popc(__ballot_sync(kFull, cnt > 0 && pre < slot_base));
        const int k = before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1;
        return run_base[max(k, 0)] + slot_base + lane;
    }
};

template <typename Key, int NSLOT, bool VARIABLE, bool SYMMETRIC, bool DENSE>
__global__ void __launch_bounds__(kQueryThreads, NSLOT <= 8 ? kQueryBlocksPerSM : 2) query_kernel(const QueryArgs<Key> a)
{
    static_assert(NSLOT % 2 == 0, "slots are processed in packed pairs");
    extern __shared__ __align__(16) int s_mem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStage st;
    st.ints = s_mem + warp * kWarpSmemInts;
    st.rec_idx = st.ints + kOffRecIdx;
    st.rec_off = st.ints + kOffRecOff;
    st.run_base = st.ints + kOffRunStart;
    float4* qbuf = reinterpret_cast<float4*>(st.ints + kOffQbuf);
    float* qr2s = reinterpret_cast<float*>(st.ints + kOffQr2);
    st.wpos = 0;
    st.nrec = 0;
    st.nb_sum = 0;

    const uint32_t hmask = (1u << a.hash_log2) - 1u;
    const unsigned lt = lanemask_lt();
    const float r2_fixed = a.r2_fixed;
    const int query_limit = a.query_limit;
    const bool same_set = a.same_set != 0;

    for (;;) {
        uint32_t c0 = 0;
        if (lane == 0) c0 = atomicAdd(a.ticket, (uint32_t)kCellsPerTicket);
        c0 = __shfl_sync(kFull, c0, 0);
        if (c0 >= (uint32_t)a.n_q_cells) break;
        const uint32_t c1 = min(c0 + (uint32_t)kCellsPerTicket, (uint32_t)a.n_q_cells);

        // The batch's cell keys and query ranges arrive with ONE coalesced load per array (lanes 0..16), cells then read them by shuffle.
        const uint32_t nb = c1 - c0;
        Key my_key = 0;
        int my_start = 0;
        if ((uint32_t)lane < nb) my_key = a.q_cell_key[c0 + lane];
        if ((uint32_t)lane <= nb) my_start = (int)a.q_cell_start[c0 + lane];

        // Software pipeline over the cells of the batch: the neighbour lookup of cell c+1 is issued before the queries of cell c are
        // processed, so its memory round trip is hidden behind that work (the lookup is ONE vector load per neighbour cell).
        Key nkey_next = 0;
        uint32_t slot_next = 0;
        bool valid_next = false;
        typename HashSlot<Key>::Raw e_next;
        uint2 d_next = make_uint2(0u, 0u);
        auto issue_lookup = [&](uint32_t i) {
            const Key key = __shfl_sync(kFull, my_key, (int)i);
            // neighbour cell offset owned by this lane (lanes 27..31 idle during the lookup).  The opaque copy of the lane id
            // keeps the per-lane Morton constants from being hoisted out of the cell loop (they would cost ~10 registers).
            int l = lane;
            asm volatile("" : "+r"(l));
            const int ox = l % 3 - 1, oy = (l / 3) % 3 - 1, oz = l / 9 - 1;
            valid_next = l < 27;
            nkey_next = morton_neighbor<Key>(key, ox, oy, oz, a.key_mask, valid_next);
            if (DENSE) {
                d_next = make_uint2(0u, 0u);
                if (valid_next) d_next = __ldg(a.dense + nkey_next);
            } else {
                slot_next = Morton<Key>::hash(nkey_next) >> (32 - a.hash_log2);
                if (valid_next) e_next = HashSlot<Key>::load(a.htable, slot_next);
            }
        };
        issue_lookup(0);

        for (uint32_t i = 0; i < nb; i++) {
            // ---------------- the 27 neighbour runs of this cell: resolve the lookup issued one iteration ago
            const int qb = __shfl_sync(kFull, my_start, (int)i), qe = __shfl_sync(kFull, my_start, (int)i + 1);
            int rs = 0, rc = 0;
            if (DENSE) {
                rs = (int)d_next.x;
                rc = (int)(d_next.y - d_next.x);
            } else if (valid_next) {
                typename HashSlot<Key>::Raw e = e_next;
                uint32_t slot = slot_next;
                for (;;) {
                    if (HashSlot<Key>::matches(e, nkey_next)) { rs = HashSlot<Key>::start(e); rc = HashSlot<Key>::count(e); break; }
                    if (HashSlot<Key>::is_empty(e)) break;
                    slot = (slot + 1) & hmask;
                    e = HashSlot<Key>::load(a.htable, slot);
                }
            }
            if (i + 1 < nb) issue_lookup(i + 1);
            const int inc = warp_inclusive_scan(rc, lane);
            const int T = __shfl_sync(kFull, inc, 31);
            RunTable runs;
            runs.pre = inc - rc;
            runs.cnt = rc;
            runs.run_base = st.run_base;
            {
                const unsigned nonempty = __ballot_sync(kFull, rc > 0);
                __syncwarp();
                if (rc > 0) st.run_base[__popc(nonempty & lt)] = rs - runs.pre;
                __syncwarp();
            }
            const int self_pre = __shfl_sync(kFull, runs.pre, 13);   // lane 13 = offset (0,0,0): the cell itself when same_set

            if (T <= NSLOT * 32) {
                // ---------------- fast path: the whole candidate list lives in registers, two slots per packed register
                f32x2 px[NSLOT / 2], py[NSLOT / 2], pz[NSLOT / 2];
                int pid[NSLOT];
                float pr2[SYMMETRIC ? NSLOT : 1];
#pragma unroll
                for (int j = 0; j < NSLOT / 2; j++) {
                    // a lane without a candidate holds a point at x = 3e38: d2 = inf, never a hit (and r2 = -1 for the symmetric test)
                    float4 v0 = make_float4(3.0e38f, 0.0f, 0.0f, __int_as_float(-1)), v1 = v0;
                    float w0 = -1.0f, w1 = -1.0f;
                    if (2 * j * 32 < T) {
                        const int pos = runs.pos(2 * j * 32, lane);
                        if (2 * j * 32 + lane < T) {
                            v0 = a.c_pts[pos];
                            if (SYMMETRIC) w0 = a.c_r2[pos];
                        }
                    }
                    if ((2 * j + 1) * 32 < T) {
                        const int pos = runs.pos((2 * j + 1) * 32, lane);
                        if ((2 * j + 1) * 32 + lane < T) {
                            v1 = a.c_pts[pos];
                            if (SYMMETRIC) w1 = a.c_r2[pos];
                        }
                    }
                    px[j] = settle2(pack2(v0.x, v1.x));
                    py[j] = settle2(pack2(v0.y, v1.y));
                    pz[j] = settle2(pack2(v0.z, v1.z));
                    pid[2 * j] = __float_as_int(v0.w);
                    pid[2 * j + 1] = __float_as_int(v1.w);
                    if (SYMMETRIC) { pr2[2 * j] = w0; pr2[2 * j + 1] = w1; }
                }
                // number of packed slot pairs in use; the query loop is instantiated per count so that its body is straight-line.
                // It returns early when the staging buffer cannot take another worst-case list; the flush lives in ONE place
                // outside the specialised loops (inlining it into each of them costs registers in the hot loop).
                // (the 16-slot variant only instantiates 2, 4, 6 and 8 pairs; unused pairs hold far-away points)
                const int npairs_exact = max((T + 63) >> 6, 1);
                const int npairs = NSLOT == 8 ? npairs_exact : min((npairs_exact + 1) & ~1, NSLOT / 2);
                auto run_queries = [&](auto np_tag, int k, const int nq, const int q0) -> int {
                    constexpr int NP = decltype(np_tag)::value;
                    for (; k < nq; k++) {
                        if (st.wpos + 1 + NP * 64 > kStageInts || st.nrec == kStageRecs) return k;
                        const float4 q = qbuf[k];
                        const int qidx = __float_as_int(q.w);
                        if (qidx >= query_limit) continue;
                        const float r2 = VARIABLE ? qr2s[k] : r2_fixed;
                        int* dst = st.ints + st.wpos + 1;
                        // candidate t = s*32 + lane is the query itself  <=>  s*32 == selfkey  (never true when selfkey < 0)
                        const int selfkey = same_set ? self_pre + (q0 + k - qb) - lane : -1;
                        const f32x2 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
                        int n = 0;
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
                                const int s = 2 * j + h;
                                bool hit = d2[h] <= r2;
                                if (SYMMETRIC) hit = hit || (d2[h] <= pr2[s]);
                                hit = hit && (selfkey != s * 32);
                                const unsigned m = __ballot_sync(kFull, hit);
                                if (hit) dst[n + __popc(m & lt)] = pid[s];
                                n += __popc(m);
                            }
                        }
                        if (lane == 0) {
                            dst[-1] = n;
                            st.rec_idx[st.nrec] = qidx;
                            st.rec_off[st.nrec] = st.wpos;
                        }
                        st.wpos += n + 1;
                        st.nrec += 1;
                    }
                    return nq;
                };
                for (int q0 = qb; q0 < qe; q0 += 32) {
                    const int qi = q0 + lane;
                    __syncwarp();
                    if (qi < qe) {
                        qbuf[lane] = a.q_pts[qi];
                        if (VARIABLE) qr2s[lane] = a.q_r2[qi];
                    }
                    __syncwarp();
                    const int nq = min(32, qe - q0);
                    int k = 0;
                    while (k < nq) {
                        if (st.wpos + 1 + npairs * 64 > kStageInts || st.nrec == kStageRecs) stage_flush(st, a, lane);
                        if (NSLOT == 8) {
                            switch (npairs) {
                            case 1: k = run_queries(std::integral_constant<int, 1>{}, k, nq, q0); break;
                            case 2: k = run_queries(std::integral_constant<int, 2>{}, k, nq, q0); break;
                            case 3: k = run_queries(std::integral_constant<int, 3>{}, k, nq, q0); break;
                            default: k = run_queries(std::integral_constant<int, 4>{}, k, nq, q0); break;
                            }
                        } else {
                            // npairs is already rounded to the instantiated counts, so the room check above and the one
                            // inside run_queries agree (otherwise this loop could spin without making progress)
                            if (npairs == 2) k = run_queries(std::integral_constant<int, 2>{}, k, nq, q0);
                            else if (npairs == 4) k = run_queries(std::integral_constant<int, 4>{}, k, nq, q0);
                            else if (npairs == 6) k = run_queries(std::integral_constant<int, 6>{}, k, nq, q0);
                            else k = run_queries(std::integral_constant<int, NSLOT / 2>{}, k, nq, q0);
                        }
                    }
                }
            } else {
                // ---------------- general path (very dense neighbourhoods): two sweeps per query, candidates re-read through L1
                for (int qi = qb; qi < qe; qi++) {
                    const float4 qv = a.q_pts[qi];
                    const int qidx = __float_as_int(qv.w);
                    if (qidx >= query_limit) continue;
                    const float r2 = VARIABLE ? a.q_r2[qi] : r2_fixed;
                    const int ts = same_set ? self_pre + (qi - qb) : -1;
                    int n = 0;
                    for (int t0 = 0; t0 < T; t0 += 32) {
                        const int t = t0 + lane;
                        const int pos = runs.pos(t0, lane);
                        bool h = false;
                        if (t < T && t != ts) {
                            const float4 v = a.c_pts[pos];
                            const float d2 = dist2(qv.x, qv.y, qv.z, v.x, v.y, v.z);
                            h = d2 <= r2;
                            if (SYMMETRIC) h = h || (d2 <= a.c_r2[pos]);
                        }
                        n += __popc(__ballot_sync(kFull, h));
                    }
                    bool ok;
                    int* dst = reserve_list(st, a, lane, qidx, n, ok);
                    if (!ok) continue;
                    int p = 1;
                    for (int t0 = 0; t0 < T; t0 += 32) {
                        const int t = t0 + lane;
                        const int pos = runs.pos(t0, lane);
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
                }
            }
        }
    }
    stage_flush(st, a, lane);
    if (lane == 0 && st.nb_sum) atomicAdd(a.n_neighbors, (unsigned long long)st.nb_sum);
}

// [min, max] of the list lengths of one pair (print_state's "n_neighbors [min, max, avg]"), computed on demand
__global__ void __launch_bounds__(256) list_minmax_kernel(const int32_t* __restrict__ ragged, const long long* __restrict__ list_pos, int n_lists, int* __restrict__ out)
{
    int lo = 0x7fffffff, hi = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lists; i += gridDim.x * blockDim.x) {
        const int n = ragged[list_pos[i]];
        lo = min(lo, n);
        hi = max(hi, n);
    }
    lo = __reduce_min_sync(kFull, lo);
    hi = __reduce_max_sync(kFull, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], lo);
        atomicMax(&out[1], hi);
    }
}

// ---- optional post pass: sort every list ascending (one warp per list, bitonic in registers for n <= 32*kSortPerLane) ----
constexpr int kListSortPerLane = 4;    // lists up to 128 ids are sorted in registers, longer ones by an in-place odd-even pass

__global__ void __launch_bounds__(256) sort_lists_kernel(int32_t* __restrict__ ragged, const long long* __restrict__ list_pos, int n_lists, int query_limit)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_lists || w >= query_limit) return;
    int32_t* l = ragged + list_pos[w];
    const int n = l[0];
    l += 1;
    if (n <= 1) return;
    if (n <= 32 * kListSortPerLane) {
        // bitonic sort over 128 virtual elements: element e lives in lane (e & 31), register (e >> 5)
        int v[kListSortPerLane];
#pragma unroll
        for (int r = 0; r < kListSortPerLane; r++) {
            const int e = r * 32 + lane;
            v[r] = e < n ? l[e] : 0x7fffffff;
        }
#pragma unroll
        for (int k = 2; k <= 32 * kListSortPerLane; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j >= 1; j >>= 1) {
                int o[kListSortPerLane];
#pragma unroll
                for (int r = 0; r < kListSortPerLane; r++) o[r] = v[r];
#pragma unroll
                for (int r = 0; r < kListSortPerLane; r++) {
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
#pragma unroll
        for (int r = 0; r < kListSortPerLane; r++) {
            const int e = r * 32 + lane;
            if (e < n) l[e] = v[r];
        }
    } else {
        // odd-even transposition in global/L2 memory (rare: > 128 neighbours)
        for (int pass = 0; pass < n; pass++) {
            for (int e = (pass & 1) + 2 * lane; e + 1 < n; e += 64) {
                const int x = l[e], y = l[e + 1];
                if (x > y) { l[e] = y; l[e + 1] = x; }
            }
            __syncwarp();
        }
    }
}

}  // namespace tnsb
