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
constexpr int kCellsPerTicket = 16;

// per-warp shared memory layout (in ints)
template <int NSLOT, bool SYMMETRIC>
struct QLayout {
    static constexpr int kStageInts = NSLOT <= 8 ? 1280 : 2048;   // staged lists [n, j0, j1, ...]; must hold one worst-case list
    static constexpr int kStageRecs = NSLOT <= 8 ? 160 : 192;     // staged list records
    static constexpr int kOffRecIdx = kStageInts;
    static constexpr int kOffRecOff = kOffRecIdx + kStageRecs;
    static constexpr int kOffRuns = kOffRecOff + kStageRecs;      // 2 parities x (base[32], pre[32], cnt[32])
    static constexpr int kOffQbuf = kOffRuns + 192;               // query float4[32]
    static constexpr int kOffQr2 = kOffQbuf + 128;                // query r^2[32]
    static constexpr int kOffCand = kOffQr2 + 32;                 // candidate float4[NSLOT*32], filled by cp.async one cell ahead
    static constexpr bool kPipe = NSLOT <= 8;                     // cp.async candidate tiles (see query_kernel)
    static constexpr int kOffCandR2 = kOffCand + (kPipe ? NSLOT * 32 * 4 : 0);  // candidate r^2[NSLOT*32] (symmetric variable radius only)
    static constexpr int kWarpInts = kOffCandR2 + (kPipe && SYMMETRIC ? NSLOT * 32 : 0);
    static constexpr int kBytes = kQueryWarps * kWarpInts * 4;
    static constexpr int kBlocksPerSM = kBytes <= 112 * 1024 ? 2 : 1;
    static_assert(kOffQbuf % 4 == 0 && kOffCand % 4 == 0 && kWarpInts % 4 == 0, "16-byte alignment of the float4 buffers");
    static_assert(kStageInts >= NSLOT * 32 + 4, "the staging buffer must hold one worst-case list");
};

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
    int bits;                     // cells per axis = 1 << bits
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
    int* ints;       // staged lists  [n, j0, j1, ...]
    int* rec_idx;    // query index of every staged list
    int* rec_off;    // its offset inside ints
    int wpos;
    int nrec;
    unsigned nb_sum; // neighbour ids written by this warp (flushed to the 64-bit global counter before it can wrap)
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
__device__ __forceinline__ int* reserve_list(WarpStage& st, const QueryArgs<Key>& a, int lane, int qidx, int n, int stage_ints, int stage_recs, bool& ok)
{
    ok = true;
    if (n + 1 > stage_ints) {
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
    if (st.wpos + n + 1 > stage_ints || st.nrec == stage_recs) stage_flush(st, a, lane);
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
// pairs live in aligned 64-bit registers (otherwise every FADD2 of the inner loop needs two extra moves for its operand).
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

// asynchronous global -> shared copies (LDGSTS): the data never passes through registers, so a warp can have the whole
// candidate tile of its NEXT cell in flight while its registers still hold the current one
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// The dense candidate list of a cell is the concatenation of its (up to 27) non-empty neighbour runs.  For the 32
// consecutive candidates of one slot, "which position of the sorted array is candidate t" is answered with one REDUX.OR,
// one VOTE and a popc per lane (no per-lane search): bit b of `starts` says that a run begins at candidate slot_base+b.
// pre / cnt: this lane's run (first candidate number, length); run_base[k] = start - pre of the k-th non-empty run.
__device__ __forceinline__ int candidate_pos(int pre, int cnt, const int* run_base, int slot_base, int lane)
{
    const bool starts_here = cnt > 0 && pre >= slot_base && pre < slot_base + 32;
    const unsigned starts = __reduce_or_sync(kFull, starts_here ? (1u << (pre - slot_base)) : 0u);
    const int before = __popc(__ballot_sync(kFull, cnt > 0 && pre < slot_base));
    const int k = before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1;
    return run_base[max(k, 0)] + slot_base + lane;
}

template <typename Key, int NSLOT, bool VARIABLE, bool SYMMETRIC, bool DENSE>
__global__ void __launch_bounds__(kQueryThreads, QLayout<NSLOT, SYMMETRIC>::kBlocksPerSM) query_kernel(const QueryArgs<Key> a)
{
    typedef QLayout<NSLOT, SYMMETRIC> LO;
    static_assert(NSLOT % 2 == 0, "slots are processed in packed pairs");
    // The cp.async tile pipeline costs a few live registers; the 16-slot variant (dense clouds, 64 candidate registers) is
    // already at the 128-register limit of 2 CTAs/SM and keeps the direct global -> register loads instead.
    constexpr bool PIPE = NSLOT <= 8;
    extern __shared__ __align__(16) int s_mem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* const wmem = s_mem + warp * LO::kWarpInts;
    WarpStage st;
    st.ints = wmem;
    st.rec_idx = wmem + LO::kOffRecIdx;
    st.rec_off = wmem + LO::kOffRecOff;
    st.wpos = 0;
    st.nrec = 0;
    st.nb_sum = 0;
    int* const runs = wmem + LO::kOffRuns;                    // [parity][base | pre | cnt][32]
    float4* const qbuf = reinterpret_cast<float4*>(wmem + LO::kOffQbuf);
    float* const qr2s = reinterpret_cast<float*>(wmem + LO::kOffQr2);
    float4* const cand = reinterpret_cast<float4*>(wmem + LO::kOffCand);
    float* const cand_r2 = reinterpret_cast<float*>(wmem + LO::kOffCandR2);

    const uint32_t hmask = (1u << a.hash_log2) - 1u;
    const unsigned lt = lanemask_lt();
    const float r2_fixed = a.r2_fixed;
    const int query_limit = a.query_limit;
    const bool same_set = a.same_set != 0;
    const uint32_t n_cells = (uint32_t)a.n_q_cells;

    // ---- neighbour lookup of one cell, split in "issue" (no waiting) and "resolve" (waits for the load issued earlier)
    Key lk_key = 0;
    uint32_t lk_slot = 0;
    bool lk_valid = false;
    typename HashSlot<Key>::Raw lk_e;
    uint2 lk_d = make_uint2(0u, 0u);
    auto issue_lookup = [&](Key cell_key) {
        // neighbour cell offset owned by this lane (lanes 27..31 idle during the lookup).  The opaque copy of the lane id
        // keeps the per-lane Morton constants from being hoisted out of the cell loop.
        int l = lane;
        asm volatile("" : "+r"(l));
        const int ox = l % 3 - 1, oy = (l / 3) % 3 - 1, oz = l / 9 - 1;
        lk_valid = l < 27;
        lk_key = morton_neighbor<Key>(cell_key, ox, oy, oz, a.key_mask, lk_valid);
        if (DENSE) {
            lk_d = make_uint2(0u, 0u);
            if (lk_valid) lk_d = __ldg(a.dense + lk_key);
        } else {
            lk_slot = Morton<Key>::hash(lk_key) >> (32 - a.hash_log2);
            if (lk_valid) lk_e = HashSlot<Key>::load(a.htable, lk_slot);
        }
    };
    // resolve the pending lookup into the run table of `parity`, start the asynchronous copy of the candidates into shared
    // memory, return T (candidate count) and the first candidate number of the cell's own run
    auto resolve_and_stage = [&](int parity, int& T_out, int& self_pre_out) {
        int rs = 0, rc = 0;
        if (DENSE) {
            rs = (int)lk_d.x;
            rc = (int)(lk_d.y - lk_d.x);
        } else if (lk_valid) {
            typename HashSlot<Key>::Raw e = lk_e;
            uint32_t slot = lk_slot;
            for (;;) {
                if (HashSlot<Key>::matches(e, lk_key)) { rs = HashSlot<Key>::start(e); rc = HashSlot<Key>::count(e); break; }
                if (HashSlot<Key>::is_empty(e)) break;
                slot = (slot + 1) & hmask;
                e = HashSlot<Key>::load(a.htable, slot);
            }
        }
        const int inc = warp_inclusive_scan(rc, lane);
        const int T = __shfl_sync(kFull, inc, 31);
        const int pre = inc - rc;
        int* rb = runs + parity * 96;
        const unsigned nonempty = __ballot_sync(kFull, rc > 0);
        if (rc > 0) rb[__popc(nonempty & lt)] = rs - pre;
        rb[32 + lane] = pre;
        rb[64 + lane] = rc;
        __syncwarp();
        if (PIPE && T <= NSLOT * 32) {
#pragma unroll
            for (int s = 0; s < NSLOT; s++) {
                if (s * 32 < T) {
                    const int pos = candidate_pos(pre, rc, rb, s * 32, lane);
                    if (s * 32 + lane < T) {
                        cp_async_16(cand + s * 32 + lane, a.c_pts + pos);
                        if (SYMMETRIC) cp_async_4(cand_r2 + s * 32 + lane, a.c_r2 + pos);
                    }
                }
            }
        }
        if (PIPE) cp_async_commit();
        T_out = T;
        self_pre_out = __shfl_sync(kFull, pre, 13);           // lane 13 = offset (0,0,0): the cell itself when same_set
    };

    // ---- batches of consecutive cells from the ticket counter; the next ticket and the next batch's keys are prefetched
    uint32_t c0 = 0;
    if (lane == 0) c0 = atomicAdd(a.ticket, (uint32_t)kCellsPerTicket);
    c0 = __shfl_sync(kFull, c0, 0);
    Key my_key = 0;
    int my_start = 0;
    if (c0 < n_cells) {
        if (c0 + lane < n_cells) my_key = a.q_cell_key[c0 + lane];
        if (c0 + lane <= n_cells) my_start = (int)a.q_cell_start[c0 + lane];
    }

    while (c0 < n_cells) {
        const int nb = (int)min((uint32_t)kCellsPerTicket, n_cells - c0);
        uint32_t t_next = 0;
        if (PIPE && lane == 0) t_next = atomicAdd(a.ticket, (uint32_t)kCellsPerTicket);     // consumed half a batch later
        Key next_key = 0;
        int next_start = 0;
        uint32_t c0_next = 0xffffffffu;

        int T_cur = 0, self_cur = 0;
        issue_lookup(__shfl_sync(kFull, my_key, 0));

        for (int i = 0; i < nb; i++) {
            // PIPE: cell i was resolved and its tile copy started one iteration ago (cell 0: here, the pipeline prologue).
            // !PIPE: every cell is resolved here; only its lookup was issued one iteration ago.
            if (!PIPE || i == 0) {
                resolve_and_stage(i & 1, T_cur, self_cur);
                if (i + 1 < nb) issue_lookup(__shfl_sync(kFull, my_key, i + 1));
            }
            const int T = T_cur, self_pre = self_cur;
            const int qb = __shfl_sync(kFull, my_start, i), qe = __shfl_sync(kFull, my_start, i + 1);
            const int* rb = runs + (i & 1) * 96;

            if (PIPE && i == (nb >> 1)) {
                // prefetch the next batch's keys / query ranges (the ticket requested at the batch start has arrived by now)
                c0_next = __shfl_sync(kFull, t_next, 0);
                if (c0_next < n_cells) {
                    if (c0_next + lane < n_cells) next_key = a.q_cell_key[c0_next + lane];
                    if (c0_next + lane <= n_cells) next_start = (int)a.q_cell_start[c0_next + lane];
                }
            }

            if (PIPE) cp_async_wait_all();
            __syncwarp();
            if (T <= NSLOT * 32) {
                // ---------------- fast path: the candidate tile moves from shared memory to registers, two slots per packed register
                f32x2 px[NSLOT / 2], py[NSLOT / 2], pz[NSLOT / 2];
                int pid[NSLOT];
                float pr2[SYMMETRIC ? NSLOT : 1];
#pragma unroll
                for (int j = 0; j < NSLOT / 2; j++) {
                    // a lane without a candidate holds a point at x = 3e38: d2 = inf, never a hit (and r2 = -1 for the symmetric test)
                    float4 v0 = make_float4(3.0e38f, 0.0f, 0.0f, __int_as_float(-1)), v1 = v0;
                    float w0 = -1.0f, w1 = -1.0f;
                    if (PIPE) {
                        if (2 * j * 32 + lane < T) {
                            v0 = cand[2 * j * 32 + lane];
                            if (SYMMETRIC) w0 = cand_r2[2 * j * 32 + lane];
                        }
                        if ((2 * j + 1) * 32 + lane < T) {
                            v1 = cand[(2 * j + 1) * 32 + lane];
                            if (SYMMETRIC) w1 = cand_r2[(2 * j + 1) * 32 + lane];
                        }
                    } else {
                        const int pre = rb[32 + lane], cnt = rb[64 + lane];
                        if (2 * j * 32 < T) {
                            const int pos = candidate_pos(pre, cnt, rb, 2 * j * 32, lane);
                            if (2 * j * 32 + lane < T) {
                                v0 = a.c_pts[pos];
                                if (SYMMETRIC) w0 = a.c_r2[pos];
                            }
                        }
                        if ((2 * j + 1) * 32 < T) {
                            const int pos = candidate_pos(pre, cnt, rb, (2 * j + 1) * 32, lane);
                            if ((2 * j + 1) * 32 + lane < T) {
                                v1 = a.c_pts[pos];
                                if (SYMMETRIC) w1 = a.c_r2[pos];
                            }
                        }
                    }
                    px[j] = settle2(pack2(v0.x, v1.x));
                    py[j] = settle2(pack2(v0.y, v1.y));
                    pz[j] = settle2(pack2(v0.z, v1.z));
                    pid[2 * j] = __float_as_int(v0.w);
                    pid[2 * j + 1] = __float_as_int(v1.w);
                    if (SYMMETRIC) { pr2[2 * j] = w0; pr2[2 * j + 1] = w1; }
                }
                if (PIPE) {
                    __syncwarp();      // every lane has read its candidates: the tile buffer may be refilled
                    // next cell: resolve its lookup, start the copy of its candidate tile, put the lookup after it in flight
                    if (i + 1 < nb) {
                        resolve_and_stage((i + 1) & 1, T_cur, self_cur);
                        if (i + 2 < nb) issue_lookup(__shfl_sync(kFull, my_key, i + 2));
                    }
                }

                // number of packed slot pairs in use; the query loop is instantiated per count so that its body is straight-line.
                // It returns early when the staging buffer cannot take another worst-case list; the flush lives in ONE place.
                const int npairs_exact = max((T + 63) >> 6, 1);
                const int npairs = NSLOT == 8 ? npairs_exact : min((npairs_exact + 1) & ~1, NSLOT / 2);
                auto run_queries = [&](auto np_tag, int k, const int nq, const int q0) -> int {
                    constexpr int NP = decltype(np_tag)::value;
                    for (; k < nq; k++) {
                        if (st.wpos + 1 + NP * 64 > LO::kStageInts || st.nrec == LO::kStageRecs) return k;
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
                        if (st.wpos + 1 + npairs * 64 > LO::kStageInts || st.nrec == LO::kStageRecs) stage_flush(st, a, lane);
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
                // ---------------- general path (very dense neighbourhoods): two sweeps per query, candidates re-read through L1/L2.
                // The run table of this cell stays valid in its parity slot while the next cell is resolved into the other one.
                if (PIPE && i + 1 < nb) {
                    resolve_and_stage((i + 1) & 1, T_cur, self_cur);
                    if (i + 2 < nb) issue_lookup(__shfl_sync(kFull, my_key, i + 2));
                }
                const int pre = rb[32 + lane], cnt = rb[64 + lane];
                for (int qi = qb; qi < qe; qi++) {
                    const float4 qv = a.q_pts[qi];
                    const int qidx = __float_as_int(qv.w);
                    if (qidx >= query_limit) continue;
                    const float r2 = VARIABLE ? a.q_r2[qi] : r2_fixed;
                    const int ts = same_set ? self_pre + (qi - qb) : -1;
                    int n = 0;
                    for (int t0 = 0; t0 < T; t0 += 32) {
                        const int t = t0 + lane;
                        const int pos = candidate_pos(pre, cnt, rb, t0, lane);
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
                    int* dst = reserve_list(st, a, lane, qidx, n, LO::kStageInts, LO::kStageRecs, ok);
                    if (!ok) continue;
                    int p = 1;
                    for (int t0 = 0; t0 < T; t0 += 32) {
                        const int t = t0 + lane;
                        const int pos = candidate_pos(pre, cnt, rb, t0, lane);
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
        // next batch (its keys were prefetched half a batch ago)
        if (PIPE) {
            c0 = c0_next;
            my_key = next_key;
            my_start = next_start;
        } else {
            // the register-tight 16-slot variant fetches the next batch synchronously
            if (lane == 0) t_next = atomicAdd(a.ticket, (uint32_t)kCellsPerTicket);
            c0 = __shfl_sync(kFull, t_next, 0);
            my_key = 0;
            my_start = 0;
            if (c0 < n_cells) {
                if (c0 + lane < n_cells) my_key = a.q_cell_key[c0 + lane];
                if (c0 + lane <= n_cells) my_start = (int)a.q_cell_start[c0 + lane];
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
