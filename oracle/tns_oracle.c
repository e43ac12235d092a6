/*
 * tns_oracle.c -- CPU restatement of the TreeNSearch neighbour criterion.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under treensearch_b200/ or include/ may
 * link, load or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may ("Oracle").
 *
 * Parity status: PINNED.  This restatement is checked against the unmodified
 * reference compiled from /root/reference (oracle/_ref/libtns_ref.so, recipe in
 * oracle/Makefile) by tests/test_oracle.py, and against the committed golden
 * fixtures under tests/golden/ that were generated from that reference.
 *
 * What is restated (citations relative to /root/reference):
 *   - neighbour criterion   TreeNSearch/source/TreeNSearch.cpp:2474-2493 (fixed radius / asymmetric)
 *                           TreeNSearch/source/TreeNSearch.cpp:2533-2554 (variable radius, symmetric)
 *                           tests/BruteforceNSearch.cpp:80-101          (same semantics, scalar)
 *   - r^2 = r*r in float    TreeNSearch/source/TreeNSearch.cpp:29, :2350-2353
 *   - self exclusion        TreeNSearch/source/TreeNSearch.cpp:2464-2466, tests/BruteforceNSearch.cpp:86
 *   - list format           TreeNSearch/source/TreeNSearch.cpp:2494-2500 ([n, j0, j1...], set-local int32 ids)
 *   - Morton bit order      TreeNSearch/extern/libmorton/morton_BMI.h:40-52 (x -> bit 0, y -> bit 1, z -> bit 2)
 *
 * The squared distance is evaluated exactly as GCC 13 contracts the reference's
 * AVX2 expression (`(dx*dx + dy*dy) + dz*dz`, -ffp-contract=fast):
 *       d2 = fmaf(dz, dz, fmaf(dx, dx, dy*dy))
 * This file must be compiled with -ffp-contract=off so that the compiler does not
 * re-associate the explicit fmaf() calls below.
 *
 * The octree of the reference (TreeNSearch.cpp:1114-1822) is NOT restated: it only
 * selects candidates and provably never drops a true neighbour.  Candidates here come
 * either from all pairs (mode 0, O(N^2)) or from a uniform grid with cell >= r_max
 * (mode 1), which is sufficient for the criterion above.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TNSO_MAX_SETS 16

typedef struct {
    int          n_sets;
    const float *pts[TNSO_MAX_SETS];
    const float *radii[TNSO_MAX_SETS];      /* per-point radii (variable radius mode) */
    int          variable[TNSO_MAX_SETS];   /* set was added through the overload with a radii array */
    int          n[TNSO_MAX_SETS];
    float        radius;                    /* fixed radius, < 0 if not set */
    int          symmetric;                 /* default 1: TreeNSearch.h:385 */
    unsigned char active[TNSO_MAX_SETS][TNSO_MAX_SETS];   /* default all 0: TreeNSearch.cpp:357-361 */
    /* results, per ordered pair */
    int64_t     *offsets[TNSO_MAX_SETS][TNSO_MAX_SETS];   /* n[set_i] + 1 */
    int32_t     *indices[TNSO_MAX_SETS][TNSO_MAX_SETS];   /* ascending within each list */
} tnso_t;

/* ------------------------------------------------------------------ criterion */

static inline float tnso_d2(const float *p, const float *q)
{
    const float dx = p[0] - q[0];
    const float dy = p[1] - q[1];
    const float dz = p[2] - q[2];
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* r2i / r2j are already squared in float. `sym` is only ever 1 in variable radius mode
 * (TreeNSearch.cpp:2431: perform_symmetric_check = !is_global_search_radius_set && symmetric_search). */
static inline int tnso_is_neighbor(float d2, float r2i, float r2j, int sym)
{
    return sym ? (d2 <= r2i || d2 <= r2j) : (d2 <= r2i);
}

/* ------------------------------------------------------------------ API */

void *tnso_create(void)
{
    tnso_t *o = (tnso_t *)calloc(1, sizeof(tnso_t));
    o->radius = -1.0f;
    o->symmetric = 1;
    return o;
}

static void tnso_free_results(tnso_t *o)
{
    for (int i = 0; i < TNSO_MAX_SETS; i++)
        for (int j = 0; j < TNSO_MAX_SETS; j++) {
            free(o->offsets[i][j]); o->offsets[i][j] = NULL;
            free(o->indices[i][j]); o->indices[i][j] = NULL;
        }
}

void tnso_destroy(void *h)
{
    tnso_t *o = (tnso_t *)h;
    tnso_free_results(o);
    free(o);
}

int tnso_add_point_set(void *h, const float *pts, const float *radii, int n, int variable)
{
    tnso_t *o = (tnso_t *)h;
    if (o->n_sets >= TNSO_MAX_SETS) return -1;
    const int s = o->n_sets++;
    o->pts[s] = pts; o->radii[s] = variable ? radii : NULL; o->n[s] = n; o->variable[s] = variable;
    return s;
}

void tnso_resize_point_set(void *h, int s, const float *pts, const float *radii, int n, int variable)
{
    tnso_t *o = (tnso_t *)h;
    o->pts[s] = pts; o->n[s] = n;
    if (variable) o->radii[s] = radii;       /* the overload without radii keeps the old pointer, TreeNSearch.cpp:97-121 */
}

void tnso_set_search_radius(void *h, float r) { ((tnso_t *)h)->radius = r; }
void tnso_set_symmetric_search(void *h, int b) { ((tnso_t *)h)->symmetric = b; }
void tnso_set_active_search(void *h, int i, int j, int b) { ((tnso_t *)h)->active[i][j] = (unsigned char)b; }

static inline float tnso_r2(const tnso_t *o, int s, int i)
{
    if (o->radius >= 0.0f) return o->radius * o->radius;       /* TreeNSearch.cpp:29 */
    const float r = o->radii[s][i];
    return r * r;                                             /* TreeNSearch.cpp:2352 */
}

static int tnso_cmp_i32(const void *a, const void *b)
{
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------ mode 0: all pairs */

static void tnso_run_bruteforce(tnso_t *o, int si, int sj)
{
    const int ni = o->n[si], nj = o->n[sj];
    const int fixed = o->radius >= 0.0f;
    const int sym = !fixed && o->symmetric;
    int64_t *off = (int64_t *)calloc((size_t)ni + 1, sizeof(int64_t));
    for (int pass = 0; pass < 2; pass++) {
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < ni; i++) {
            const float *p = o->pts[si] + 3 * (size_t)i;
            const float r2i = tnso_r2(o, si, i);
            int64_t c = 0;
            int32_t *dst = pass ? o->indices[si][sj] + off[i] : NULL;
            for (int j = 0; j < nj; j++) {
                if (si == sj && i == j) continue;
                const float d2 = tnso_d2(p, o->pts[sj] + 3 * (size_t)j);
                const float r2j = sym ? tnso_r2(o, sj, j) : 0.0f;
                if (tnso_is_neighbor(d2, r2i, r2j, sym)) {
                    if (pass) dst[c] = j;
                    c++;
                }
            }
            if (!pass) off[i + 1] = c;
        }
        if (!pass) {
            for (int i = 0; i < ni; i++) off[i + 1] += off[i];
            o->indices[si][sj] = (int32_t *)malloc(sizeof(int32_t) * (size_t)(off[ni] > 0 ? off[ni] : 1));
        }
    }
    o->offsets[si][sj] = off;
}

/* ------------------------------------------------------------------ mode 1: uniform grid */

typedef struct {
    int      dim[3];
    double   bottom[3];
    double   inv;
    int64_t *cell_start;     /* n_cells + 1 */
    int32_t *order;          /* point ids sorted by cell, ascending id inside a cell */
} tnso_grid_t;

static inline int tnso_cell_coord(const tnso_grid_t *g, float x, int d)
{
    int c = (int)floor(((double)x - g->bottom[d]) * g->inv);
    if (c < 0) c = 0;
    if (c >= g->dim[d]) c = g->dim[d] - 1;
    return c;
}

static void tnso_grid_build(tnso_grid_t *g, const float *pts, int n, const double bottom[3], const double top[3], double cell)
{
    g->inv = 1.0 / cell;
    int64_t n_cells = 1;
    for (int d = 0; d < 3; d++) {
        g->bottom[d] = bottom[d];
        g->dim[d] = (int)floor((top[d] - bottom[d]) * g->inv) + 1;
        n_cells *= g->dim[d];
    }
    g->cell_start = (int64_t *)calloc((size_t)n_cells + 1, sizeof(int64_t));
    g->order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int64_t *cell_of = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        const float *p = pts + 3 * (size_t)i;
        const int64_t c = ((int64_t)tnso_cell_coord(g, p[2], 2) * g->dim[1] + tnso_cell_coord(g, p[1], 1)) * g->dim[0] + tnso_cell_coord(g, p[0], 0);
        cell_of[i] = c;
        g->cell_start[c + 1]++;
    }
    for (int64_t c = 0; c < n_cells; c++) g->cell_start[c + 1] += g->cell_start[c];
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_cells > 0 ? n_cells : 1));
    memcpy(cursor, g->cell_start, sizeof(int64_t) * (size_t)n_cells);
    for (int i = 0; i < n; i++) g->order[cursor[cell_of[i]]++] = i;   /* stable: ascending id per cell */
    free(cursor);
    free(cell_of);
}

static void tnso_grid_free(tnso_grid_t *g) { free(g->cell_start); free(g->order); }

static void tnso_run_grid(tnso_t *o)
{
    const int fixed = o->radius >= 0.0f;
    const int sym = !fixed && o->symmetric;

    /* world box + largest radius over all sets */
    double bottom[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, top[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    double r_max = fixed ? (double)o->radius : 0.0;
    int64_t n_total = 0;
    for (int s = 0; s < o->n_sets; s++) {
        n_total += o->n[s];
        for (int i = 0; i < o->n[s]; i++) {
            for (int d = 0; d < 3; d++) {
                const double v = o->pts[s][3 * (size_t)i + d];
                if (v < bottom[d]) bottom[d] = v;
                if (v > top[d]) top[d] = v;
            }
            if (!fixed && o->radii[s][i] > r_max) r_max = o->radii[s][i];
        }
    }
    if (n_total == 0) { for (int d = 0; d < 3; d++) { bottom[d] = 0; top[d] = 0; } }
    /* cell edge strictly larger than any search distance (0.1% margin absorbs float rounding of d2);
       coarsen until the dense table stays small */
    double cell = r_max * 1.001;
    if (!(cell > 0.0)) cell = 1.0;
    for (;;) {
        double cells = 1.0;
        for (int d = 0; d < 3; d++) cells *= floor((top[d] - bottom[d]) / cell) + 1.0;
        if (cells <= 4.0e8) break;
        cell *= 2.0;
    }

    tnso_grid_t grids[TNSO_MAX_SETS];
    for (int s = 0; s < o->n_sets; s++) tnso_grid_build(&grids[s], o->pts[s], o->n[s], bottom, top, cell);

    for (int si = 0; si < o->n_sets; si++) {
        for (int sj = 0; sj < o->n_sets; sj++) {
            if (!o->active[si][sj]) continue;
            const tnso_grid_t *g = &grids[sj];
            const int ni = o->n[si];
            int64_t *off = (int64_t *)calloc((size_t)ni + 1, sizeof(int64_t));
            for (int pass = 0; pass < 2; pass++) {
                #pragma omp parallel for schedule(dynamic, 1024)
                for (int i = 0; i < ni; i++) {
                    const float *p = o->pts[si] + 3 * (size_t)i;
                    const float r2i = tnso_r2(o, si, i);
                    const int cx = tnso_cell_coord(g, p[0], 0), cy = tnso_cell_coord(g, p[1], 1), cz = tnso_cell_coord(g, p[2], 2);
                    int64_t c = 0;
                    int32_t *dst = pass ? o->indices[si][sj] + off[i] : NULL;
                    for (int z = (cz > 0 ? cz - 1 : 0); z <= cz + 1 && z < g->dim[2]; z++)
                    for (int y = (cy > 0 ? cy - 1 : 0); y <= cy + 1 && y < g->dim[1]; y++) {
                        const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx + 1 < g->dim[0] ? cx + 1 : g->dim[0] - 1;
                        const int64_t row = ((int64_t)z * g->dim[1] + y) * g->dim[0];
                        for (int64_t k = g->cell_start[row + x0]; k < g->cell_start[row + x1 + 1]; k++) {
                            const int j = g->order[k];
                            if (si == sj && i == j) continue;
                            const float d2 = tnso_d2(p, o->pts[sj] + 3 * (size_t)j);
                            const float r2j = sym ? tnso_r2(o, sj, j) : 0.0f;
                            if (tnso_is_neighbor(d2, r2i, r2j, sym)) {
                                if (pass) dst[c] = j;
                                c++;
                            }
                        }
                    }
                    if (!pass) off[i + 1] = c;
                    else if (c > 1) qsort(dst, (size_t)c, sizeof(int32_t), tnso_cmp_i32);
                }
                if (!pass) {
                    for (int i = 0; i < ni; i++) off[i + 1] += off[i];
                    o->indices[si][sj] = (int32_t *)malloc(sizeof(int32_t) * (size_t)(off[ni] > 0 ? off[ni] : 1));
                }
            }
            o->offsets[si][sj] = off;
        }
    }
    for (int s = 0; s < o->n_sets; s++) tnso_grid_free(&grids[s]);
}

/* mode 0 = all pairs (tests/BruteforceNSearch.cpp:66-105), mode 1 = uniform grid candidates */
int tnso_run(void *h, int mode)
{
    tnso_t *o = (tnso_t *)h;
    tnso_free_results(o);
    const int fixed = o->radius >= 0.0f;
    for (int s = 0; s < o->n_sets; s++) {
        if (fixed && o->variable[s]) return -1;           /* TreeNSearch.cpp:383-386 */
        if (!fixed && !o->variable[s]) return -2;         /* TreeNSearch.cpp:388-391 */
    }
    if (mode == 0) {
        for (int si = 0; si < o->n_sets; si++)
            for (int sj = 0; sj < o->n_sets; sj++)
                if (o->active[si][sj]) tnso_run_bruteforce(o, si, sj);
    } else {
        tnso_run_grid(o);
    }
    return 0;
}

int64_t tnso_pair_total(void *h, int si, int sj)
{
    tnso_t *o = (tnso_t *)h;
    return o->offsets[si][sj] ? o->offsets[si][sj][o->n[si]] : -1;
}

/* copies CSR of pair (si, sj): offsets[n_i + 1], indices[total] (ascending per list) */
int tnso_pair_export(void *h, int si, int sj, int64_t *offsets, int32_t *indices)
{
    tnso_t *o = (tnso_t *)h;
    if (!o->offsets[si][sj]) return -1;
    const int ni = o->n[si];
    memcpy(offsets, o->offsets[si][sj], sizeof(int64_t) * ((size_t)ni + 1));
    memcpy(indices, o->indices[si][sj], sizeof(int32_t) * (size_t)o->offsets[si][sj][ni]);
    return 0;
}

/* ------------------------------------------------------------------ helpers shared with the tests */

/* libmorton bit order (morton_BMI.h:40-52): x occupies bits 0,3,6..., y bits 1,4,7..., z bits 2,5,8... */
uint64_t tnso_morton3d_64(uint32_t x, uint32_t y, uint32_t z)
{
    uint64_t k = 0;
    for (int b = 0; b < 21; b++) {
        k |= ((uint64_t)((x >> b) & 1u)) << (3 * b);
        k |= ((uint64_t)((y >> b) & 1u)) << (3 * b + 1);
        k |= ((uint64_t)((z >> b) & 1u)) << (3 * b + 2);
    }
    return k;
}

/* Order independent 64-bit digest of every neighbour list of a CSR: out[i] = sum_j mix(j) ; used to compare
 * 10M-point results without sorting lists. `pos` gives the first id of list i, `cnt` its length. */
static inline uint64_t tnso_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void tnso_list_digests(const int32_t *ids, const int64_t *pos, const int32_t *cnt, int64_t n, uint64_t *out)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        uint64_t acc = tnso_mix64((uint64_t)cnt[i] ^ 0xABCDull);
        const int32_t *l = ids + pos[i];
        for (int32_t k = 0; k < cnt[i]; k++) acc += tnso_mix64((uint64_t)(uint32_t)l[k]);
        out[i] = acc;
    }
}
