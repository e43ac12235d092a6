/*
 * tnsb.h -- C ABI of the B200-native fixed-radius neighbour search engine (libtnsb.so).
 *
 * This is the drop-in boundary for the hot path of InteractiveComputerGraphics/TreeNSearch:
 * everything `tns::TreeNSearch::run()` / `prepare_zsort()` do (reference: TreeNSearch/source/TreeNSearch.cpp:138-149,
 * :2571-2662) happens behind these entry points in hand-written sm_100a CUDA.  The C++ class
 * `tns::TreeNSearch` in include/tns/TreeNSearch.h and the Python mirror treensearch_b200.TreeNSearch are thin
 * forwards to this ABI.  Plain pointers and sizes only: no STL, CUDA or torch types cross it.
 *
 * Conventions
 *   - every function that can fail returns an int status: TNSB_OK (0) or a negative TNSB_ERR_* code; the message is
 *     available from tnsb_last_error().  The reference prints the message to std::cout and calls exit(-1)
 *     (TreeNSearch.cpp:22-25, :366-392, :510-515); the C++ shim reproduces that on top of these codes.
 *   - point / radii pointers are BORROWED, exactly as in the reference (TreeNSearch.cpp:35-41): they are re-read on every
 *     tnsb_run() and must stay valid until replaced with tnsb_resize_point_set_*.  They may point to pageable host memory,
 *     pinned host memory or device memory of the context's GPU (detected per call with cudaPointerGetAttributes).
 *   - neighbour ids are set-local int32 indices into set_j (TreeNSearch.cpp:2015, :2260).
 */
#ifndef TNSB_H
#define TNSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnsb_context tnsb_context;

enum {
    TNSB_OK = 0,
    TNSB_ERR_INVALID_ARGUMENT = -1,   /* bad set id, null handle, ... (reference: assert / exit(-1)) */
    TNSB_ERR_INVALID_STATE = -2,      /* the configuration errors of TreeNSearch::_check(), TreeNSearch.cpp:366-392 */
    TNSB_ERR_CUDA = -3,               /* a CUDA runtime call failed (message holds cudaGetErrorString) */
    TNSB_ERR_LIMIT = -4,              /* a documented size limit was exceeded */
    TNSB_ERR_NO_DEVICE = -5           /* no usable CUDA device: the engine has no CPU fallback */
};

/* options for tnsb_set_option() */
enum {
    TNSB_OPT_HOST_RESULTS = 1,   /* 1 (default): mirror neighbour lists into pinned host memory at the end of tnsb_run();
                                    0: keep them in HBM only (tnsb_get_neighborlists_device) */
    TNSB_OPT_PIN_USER_MEMORY = 2,/* 1: cudaHostRegister borrowed pageable arrays once and cache the registration; 0 (default): stage
                                    pageable arrays through an internal pinned buffer */
    TNSB_OPT_LIST_CAPACITY = 3,  /* initial capacity, in ints per searching point, of the ragged list buffer (default 48) */
    TNSB_OPT_QUERY_LIMIT = 4,    /* >= 0: in every set only points with index < value are searching points; the remaining
                                    points are find-only ("ghost"/halo points of a Z-slab shard).  -1 (default): all points search */
    TNSB_OPT_SORT_LISTS = 5,     /* neighbour ids ascending inside every list, like the reference's lists: 1 = always, 0 = never
                                    (cell-traversal order), -1 (default) = whenever the lists are mirrored to the host: the brick query ranks
                                    every list in shared memory before it leaves the SM, which hides completely under the PCIe writes */
    TNSB_OPT_ZERO_COPY_RESULTS = 7, /* with HOST_RESULTS: 1 (default) = the query kernel writes the lists straight into mapped pinned host
                                    memory (PCIe writes overlap the search, no HBM copy of the ids, no D2H afterwards; measured 28.1 ms
                                    vs 29.7 ms end to end at 10M points); 0 = lists in HBM, then one D2H copy.  Ignored (HBM path) when
                                    TNSB_OPT_SORT_LISTS is set or HOST_RESULTS is 0 */
    TNSB_OPT_QUERY_KERNEL = 8,   /* which distance query runs: 0 (default) = automatic: the brick query (half-radius grid with linear row keys + prefix
                                    cell table, slab rows staged by TMA bulk copies, a lane owns a QUERY; csrc/query_brick.cuh) while the cell table is
                                    affordable (cells <= max(2^22, 8 x points)), else the cell kernel; 1 = always the cell kernel (cell = radius, 3-D
                                    Morton keys, dense table or hash of the occupied cells, a lane owns a CANDIDATE, ballot compaction; csrc/query.cuh:
                                    huge sparse domains, 64-bit keys).  Same neighbour sets.  TNSB_QUERY_KERNEL=0|1 in the environment sets the
                                    default of new contexts */
    TNSB_OPT_BUILD = 9,          /* how the sorted grid is built: 0 (default) = automatic: a bucket build (ONE counting pass over the full cell
                                    key: cell populations by L2 atomics, exclusive scan, scatter of the (x, y, z, id) records) while the cell
                                    table is small next to the point count, else the LSD radix sort of (key, index) pairs; 1 = always the
                                    radix sort (stable: points of a cell keep their input order; the bucket build leaves them in arrival
                                    order, which changes the order INSIDE neighbour lists, never the sets).  TNSB_BUILD=0|1 in the
                                    environment sets the default of new contexts */
    TNSB_OPT_POINT_STRIDE = 6    /* floats between consecutive points of float32 sets: 3 (default, xyzxyz as in the reference) or
                                    4 ((x, y, z, id) records as produced by tnsb_shard_partition; the 4th word is ignored) */
};

/* timings (milliseconds, CUDA events on the engine's stream) and sizes of the last tnsb_run() */
typedef struct tnsb_stats {
    double ms_total_device;      /* first kernel start .. last kernel end (excludes H2D / D2H) */
    double ms_upload;            /* host -> device copies of points / radii */
    double ms_aabb;              /* world box + radius reduction */
    double ms_keys;              /* cell assignment + Morton keys */
    double ms_sort;              /* radix sort of (key, index) */
    double ms_reorder;           /* gather of points into sorted order */
    double ms_cells;             /* cell start/end compaction + cell hash table */
    double ms_query;             /* 27-cell query kernels, all active pairs */
    double ms_download;          /* device -> host copy of lists */
    double ms_wall;              /* host wall clock of the whole call */
    int64_t n_points_total;
    int64_t n_queries;           /* sum over active pairs of searching points */
    int64_t n_neighbors;         /* sum over active pairs of neighbour ids written */
    int64_t n_list_ints;         /* ints in the ragged buffers (ids + one count word per list) */
    int64_t n_cells;             /* occupied cells, all sets */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int32_t n_kernel_launches;   /* kernels launched by the last tnsb_run() */
    int32_t n_query_launches;
    int32_t key_bits;            /* Morton key width used by the sort */
    int32_t sort_passes;
    int32_t n_reruns;            /* query re-launches after a list buffer overflow */
    float   cell_size;           /* grid cell edge actually used (>= largest search radius) */
    float   domain_bottom[3];
    float   domain_top[3];
    int32_t brick_query;         /* 1: the last run used the brick query (half-radius grid), 0: the cell kernel */
    int64_t n_slow_queries;      /* brick query: queries answered by its warp-cooperative slow path (dense cells, long lists) */
    int32_t max_list;            /* brick query: longest neighbour list of the last run (1000: some list overflowed its column) */
    int32_t speculative_grid;    /* 1: the last run reused the previous run's grid without waiting for the world box (checked on the device) */
    int32_t graph_replay;        /* 1: the enqueue phase of the last run was ONE CUDA graph launch (small problems in steady state; per-stage timings are 0,
                                    ms_total_device spans the whole graph).  TNSB_GRAPH=0 in the environment disables it */
} tnsb_stats;

/* ---- life cycle --------------------------------------------------------------------------------------------------- */
/* replaces: tns::TreeNSearch::TreeNSearch()  (TreeNSearch.h:37).  device < 0 selects the current CUDA device. */
int  tnsb_create(tnsb_context** out, int device);
/* replaces: tns::TreeNSearch::~TreeNSearch() (TreeNSearch.h:38) */
void tnsb_destroy(tnsb_context* ctx);
/* message of the last failing call on this context ("" if none); ctx == NULL returns the message of a failed tnsb_create */
const char* tnsb_last_error(const tnsb_context* ctx);
/* library version string */
const char* tnsb_version(void);

/* ---- point sets --------------------------------------------------------------------------------------------------- */
/* replaces: add_point_set(const float*, int) and (const float*, const float*, int)   TreeNSearch.h:50,112 / .cpp:35-41,49-57
   variable_radius = 1 stands for the overloads that take a radii array (the reference tells the modes apart by overload, not by
   pointer value: an empty set is added as (nullptr, nullptr, 0), tests/tests.cpp:369); with variable_radius = 0 `radii` is ignored.
   Returns the new set id (>= 0) or a negative error. */
int tnsb_add_point_set_f32(tnsb_context* ctx, const float* points_xyz, const float* radii, int n_points, int variable_radius);
/* replaces: add_point_set(const double*, int) and (const double*, const double*, int) TreeNSearch.h:63,126 / .cpp:42-48,58-66
   doubles are converted with (float) per component on the device (reference: TreeNSearch.cpp:275-296). */
int tnsb_add_point_set_f64(tnsb_context* ctx, const double* points_xyz, const double* radii, int n_points, int variable_radius);
/* replaces: resize_point_set(...) x4   TreeNSearch.h:72,81,136,146 / .cpp:67-133
   variable_radius = 0 (the overloads without radii) leaves the set's radii pointer untouched, like the reference. */
int tnsb_resize_point_set_f32(tnsb_context* ctx, int set_id, const float* points_xyz, const float* radii, int n_points, int variable_radius);
int tnsb_resize_point_set_f64(tnsb_context* ctx, int set_id, const double* points_xyz, const double* radii, int n_points, int variable_radius);

/* ---- search configuration ----------------------------------------------------------------------------------------- */
/* replaces: set_search_radius(float|double)  TreeNSearch.h:90,99 / .cpp:20-34 */
int tnsb_set_search_radius(tnsb_context* ctx, float radius);
/* replaces: set_cell_size(float|double)      TreeNSearch.h:156,166 / .cpp:135-137,173-182.  Kept for API parity (may be
   called once, like the reference); the engine's grid cell is always >= the largest search radius, see DESIGN.md. */
int tnsb_set_cell_size(tnsb_context* ctx, float cell_size);
/* replaces: set_symmetric_search(bool)       TreeNSearch.h:225 / .cpp:169-172 */
int tnsb_set_symmetric_search(tnsb_context* ctx, int active);
/* replaces: set_active_search(int,int,bool)  TreeNSearch.h:265 / .cpp:221-224 */
int tnsb_set_active_search(tnsb_context* ctx, int set_i, int set_j, int active);
/* replaces: set_active_search(int,bool,bool) TreeNSearch.h:275 / .cpp:225-235 (find column first, then search row) */
int tnsb_set_active_search_of_set(tnsb_context* ctx, int set_i, int search_neighbors, int find_neighbors);
/* replaces: set_all_searches(bool)           TreeNSearch.h:256 / .cpp:236-243 */
int tnsb_set_all_searches(tnsb_context* ctx, int active);
/* engine options (TNSB_OPT_*) */
int tnsb_set_option(tnsb_context* ctx, int option, int64_t value);
/* run every kernel and copy of this context on the caller's CUDA stream (a cudaStream_t passed as void*; NULL is the legacy
   default stream, (void*)-1 restores the context's own non-blocking stream).  Lets a host framework order the search against
   its own work (e.g. NCCL collectives that produced the points) and time it with its own events. */
int tnsb_set_stream(tnsb_context* ctx, void* cuda_stream);

/* ---- getters (TreeNSearch.h:304-334 / .cpp:191-220) --------------------------------------------------------------- */
int tnsb_get_n_sets(const tnsb_context* ctx);
int tnsb_get_n_points_in_set(const tnsb_context* ctx, int set_i);
int tnsb_get_total_n_points(const tnsb_context* ctx);
int tnsb_is_search_active(const tnsb_context* ctx, int set_i, int set_j);
int tnsb_does_set_exist(const tnsb_context* ctx, int set_i);

/* ---- the hot path --------------------------------------------------------------------------------------------------- */
/* replaces: tns::TreeNSearch::run() (and run_scalar())   TreeNSearch.h:171,233 / .cpp:138-160
   upload -> world box -> cell hash + Morton keys -> radix sort (key,index) -> reorder -> cell start/end -> 27-cell query
   -> ragged neighbour lists (-> pinned host mirror). */
int tnsb_run(tnsb_context* ctx);

/* replaces: get_neighborlist(set_i,set_j,i)   TreeNSearch.h:182 / .cpp:241-249 and NeighborList (NeighborList.h:8-39)
   Host view of the result of pair (set_i -> set_j):  list of point i = ragged + list_pos[i], laid out exactly like the
   reference's storage  [n, j0, j1, ..., j(n-1)]  (TreeNSearch.h:395).  Pointers stay valid until the next tnsb_run(). */
int tnsb_get_neighborlists(const tnsb_context* ctx, int set_i, int set_j,
                           const int32_t** ragged, const int64_t** list_pos, int64_t* n_ints);
/* same host view with 32-bit positions: list_pos travels over PCIe as uint32 while the list buffer holds fewer than 2^32 ints (half the
   bytes; tnsb_get_neighborlists then widens it on the host the first time it is asked).  TNSB_ERR_LIMIT when the positions need 64 bits.
   This is what include/tns/TreeNSearch.h reads (TreeNSearch.h:395 of the reference keeps one pointer per point instead). */
int tnsb_get_neighborlists_u32(const tnsb_context* ctx, int set_i, int set_j,
                               const int32_t** ragged, const uint32_t** list_pos32, int64_t* n_ints);
/* same, device pointers (always available after tnsb_run(), also when TNSB_OPT_HOST_RESULTS == 0) */
int tnsb_get_neighborlists_device(const tnsb_context* ctx, int set_i, int set_j,
                                  const int32_t** d_ragged, const int64_t** d_list_pos, int64_t* n_ints);

/* replaces: prepare_zsort()     TreeNSearch.h:202 / .cpp:2571-2716.  new -> old permutation per set by Morton key of the grid cell
   (x lowest bit, libmorton order), stable inside a cell.  After a tnsb_run() whose grid is still valid (no resize since) the order
   comes from the records resident in HBM (no upload, like the reference's reuse of its cells, :2598-2661). */
int tnsb_prepare_zsort(tnsb_context* ctx);
/* replaces: get_zsort_order(set) TreeNSearch.h:334 / .cpp:250-253 */
int tnsb_get_zsort_order(const tnsb_context* ctx, int set_i, const int32_t** new_to_old, int* n_points);
/* replaces: apply_zsort<T>(set_i, T*, stride)  TreeNSearch.h:443-481, for arrays resident in HBM: ONE launch gathers up to 8 arrays of
   set_i (positions, velocities, ...): dst_k[new] = src_k[old], rows of row_bytes[k] bytes (a multiple of 4: any T / stride the
   reference's template takes).  dst_k == src_k permutes in place through an internal staging copy, like the reference does.
   (host arrays of arbitrary T are gathered by the header template) */
int tnsb_apply_zsort_device(tnsb_context* ctx, int set_i, int n_arrays, const void* const* d_src, void* const* d_dst, const int* row_bytes);
/* the same for one float32 array in place: data[new*stride + c] = tmp[old*stride + c] */
int tnsb_apply_zsort_device_f32(tnsb_context* ctx, int set_i, float* d_data, int stride);

/* ---- multi-GPU: Z-slab decomposition helpers (no counterpart in the single-process reference) ---------------- */
/* min/max of a device-resident chunk of points: out = {min x, y, z, max x, y, z}.  Synchronises the stream. */
int tnsb_shard_aabb(tnsb_context* ctx, const float* d_points, int n_points, int stride, float out_min_max[6]);
/* histogram (n_bins <= 8192 uint32 bins over [lo, hi)) of one coordinate of a device-resident chunk; asynchronous on the stream.
   All-reduced over the ranks it yields slab cuts with equal point counts. */
int tnsb_shard_histogram(tnsb_context* ctx, const float* d_points, int n_points, int stride, int axis, float lo, float hi,
                         int n_bins, uint32_t* d_hist);
/* Buckets a device-resident chunk into (x, y, z, bits(id_base + i)) float4 records ordered
       [owned by part 0 | ... | owned by part P-1 | halo of part 0 | ... | halo of part P-1]
   part g owns coordinates in [cuts[g], cuts[g+1]) (cuts[0] / cuts[P] are ignored: -inf / +inf) and gets as halo every other
   point within `halo` of that interval.  counts_out[0..P) = owned counts, counts_out[P..2P) = halo counts.  Synchronises. */
int tnsb_shard_partition(tnsb_context* ctx, const float* d_points, int n_points, int stride, int id_base, int axis,
                         const float* cuts, int n_parts, float halo, float* d_records, int64_t capacity_records, int64_t* counts_out);

/* One-sided exchange over peer memory (NVLink / NVSwitch): partition + exchange in ONE kernel, no count pass, no all-to-all.
   Every rank (one process per GPU) creates two receive windows in its HBM (step parity), publishes their CUDA IPC handles
   (2 x 64 bytes; exchange them with any host side all-gather) and opens the windows of all ranks.  Per step:
       tnsb_shard_push(parity)    routes every local point into its owner's window and into the windows that need it as halo
       <barrier across the ranks, stream ordered behind the push: e.g. a 4-byte NCCL all_reduce(MAX) of d_flag>
       tnsb_shard_collect(parity) -> device pointer to this rank's [owned | halo] (x, y, z, bits(global id)) records and their counts;
                                     feed it to tnsb_resize_point_set_f32 with TNSB_OPT_POINT_STRIDE = 4, TNSB_OPT_QUERY_LIMIT = n_owned.
   Alternate parity 0 / 1 from step to step: peers may already push step k+1 while this rank still searches step k.
   d_flag >= 2 after the barrier (and TNSB_ERR_LIMIT from the owner's collect): a window was too small -- the counts returned by
   collect are exact: all ranks create larger windows and repeat the step. */
int tnsb_shard_window_create(tnsb_context* ctx, int64_t capacity_owned_records, int64_t capacity_halo_records, unsigned char* ipc_handles_out /* 128 bytes */);
int tnsb_shard_window_open(tnsb_context* ctx, int n_ranks, int my_rank, const unsigned char* all_ipc_handles /* n_ranks x 128 bytes, rank major */);
int tnsb_shard_push(tnsb_context* ctx, int parity, const float* d_points, int n_points, int stride, int id_base, int axis,
                    const float* cuts, int n_parts, float halo, int* d_flag /* device int, raised to 2 when a window overflows; may be NULL */);
int tnsb_shard_collect(tnsb_context* ctx, int parity, float** d_records, int64_t* n_owned, int64_t* n_halo);
/* the same, and the value of the barrier's device flag (d_flag after the all_reduce) comes back in the same host round trip */
int tnsb_shard_collect_flag(tnsb_context* ctx, int parity, float** d_records, int64_t* n_owned, int64_t* n_halo, const int* d_flag, int* flag_out);

/* ---- diagnostics ---------------------------------------------------------------------------------------------------- */
/* replaces: get_neighborlist_n_bytes()  TreeNSearch.h:246 / .cpp:254-261 */
uint64_t tnsb_get_neighborlist_n_bytes(const tnsb_context* ctx);
/* timings / sizes of the last run (feeds print_state(), TreeNSearch.cpp:2718-2873, and bench.py) */
int tnsb_get_stats(const tnsb_context* ctx, tnsb_stats* out);
/* per pair [min, max, sum] of the neighbour counts of the last run (print_state's "n_neighbors set_i -> set_j") */
int tnsb_get_pair_neighbor_stats(const tnsb_context* ctx, int set_i, int set_j, int64_t out_min_max_sum[3]);

#ifdef __cplusplus
}
#endif
#endif /* TNSB_H */
