/*
 * ref_shim.cpp -- C-ABI wrapper around the UNMODIFIED reference (tns::TreeNSearch and the
 * tests' BruteforceNSearch), compiled from the sources where they lie under /root/reference
 * into oracle/_ref/libtns_ref.so by oracle/Makefile.  No reference source is copied into this
 * repository; this file only *includes* the reference's public headers at build time.
 *
 * TEST INFRASTRUCTURE ONLY: used to pin the restated oracle (oracle/tns_oracle.c), to generate
 * the golden fixtures (tests/golden/make_golden.py) and as the timed CPU baseline
 * (bench.py cpu_baseline / --impl reference).  The product never loads it.
 *
 * mode 0 = tns::TreeNSearch::run()         (AVX2 path,  TreeNSearch.cpp:138-149)  <- "the reference result"
 * mode 1 = tns::TreeNSearch::run_scalar()  (scalar,     TreeNSearch.cpp:150-160)
 * mode 2 = BruteforceNSearch::run()        (tests/BruteforceNSearch.cpp:66-105)
 */
#include <TreeNSearch>
#include "BruteforceNSearch.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace {
struct RefCtx {
    tns::TreeNSearch tns;
    BruteforceNSearch bf;
    std::vector<int> n;
    std::vector<const float*> pts;
    std::vector<const float*> radii;
    std::vector<int> variable;
    float radius = -1.0f;
    int last_mode = 0;
    bool bf_built = false;
};
}

extern "C" {

void* tnsref_create() { return new RefCtx(); }
void tnsref_destroy(void* h) { delete static_cast<RefCtx*>(h); }

int tnsref_n_threads() { return omp_get_max_threads(); }

void tnsref_set_search_radius(void* h, float r)
{
    auto* c = static_cast<RefCtx*>(h);
    c->radius = r;
    c->tns.set_search_radius(r);
}

/* `variable` selects the overload with a radii array (an empty variable-radius set is (nullptr, nullptr, 0), tests/tests.cpp:369) */
int tnsref_add_point_set(void* h, const float* pts, const float* radii, int n, int variable)
{
    auto* c = static_cast<RefCtx*>(h);
    c->n.push_back(n); c->pts.push_back(pts); c->radii.push_back(radii); c->variable.push_back(variable);
    return variable ? c->tns.add_point_set(pts, radii, n) : c->tns.add_point_set(pts, n);
}

int tnsref_add_point_set_f64(void* h, const double* pts, const double* radii, int n, int variable)
{
    auto* c = static_cast<RefCtx*>(h);
    c->n.push_back(n); c->pts.push_back(nullptr); c->radii.push_back(nullptr); c->variable.push_back(variable);
    return variable ? c->tns.add_point_set(pts, radii, n) : c->tns.add_point_set(pts, n);
}

void tnsref_resize_point_set(void* h, int s, const float* pts, const float* radii, int n, int variable)
{
    auto* c = static_cast<RefCtx*>(h);
    c->n[s] = n; c->pts[s] = pts;
    if (variable) { c->radii[s] = radii; c->tns.resize_point_set(s, pts, radii, n); }
    else c->tns.resize_point_set(s, pts, n);
}

void tnsref_set_active_search(void* h, int i, int j, int b) { static_cast<RefCtx*>(h)->tns.set_active_search(i, j, b != 0); }
void tnsref_set_symmetric_search(void* h, int b) { static_cast<RefCtx*>(h)->tns.set_symmetric_search(b != 0); }
void tnsref_set_n_threads(void* h, int n) { static_cast<RefCtx*>(h)->tns.set_n_threads(n); }

static void build_bf(RefCtx* c)
{
    // a fresh BruteforceNSearch mirroring the TreeNSearch configuration (float sets only)
    c->bf = BruteforceNSearch();
    const int ns = c->tns.get_n_sets();
    for (int s = 0; s < ns; s++) {
        if (c->variable[s]) c->bf.add_point_set(c->pts[s], c->radii[s], c->n[s]);
        else                c->bf.add_point_set(c->pts[s], c->radius, c->n[s]);
    }
    for (int i = 0; i < ns; i++)
        for (int j = 0; j < ns; j++)
            c->bf.set_active_search(i, j, c->tns.is_search_active(i, j));
}

int tnsref_run(void* h, int mode)
{
    auto* c = static_cast<RefCtx*>(h);
    c->last_mode = mode;
    if (mode == 0) c->tns.run();
    else if (mode == 1) c->tns.run_scalar();
    else { build_bf(c); c->bf.run(); }
    return 0;
}

/* symmetric flag for the brute force arm must be passed explicitly (the reference offers no getter) */
void tnsref_bf_set_symmetric(void* h, int b) { static_cast<RefCtx*>(h)->bf.set_symmetric_search(b != 0); }

int tnsref_run_bruteforce(void* h, int symmetric)
{
    auto* c = static_cast<RefCtx*>(h);
    c->last_mode = 2;
    build_bf(c);
    c->bf.set_symmetric_search(symmetric != 0);
    c->bf.run();
    return 0;
}

int64_t tnsref_pair_total(void* h, int si, int sj)
{
    auto* c = static_cast<RefCtx*>(h);
    const int ni = c->n[si];
    int64_t total = 0;
    if (c->last_mode == 2) {
        for (int i = 0; i < ni; i++) total += c->bf.get_n_neighbors(si, sj, i);
    } else {
        #pragma omp parallel for reduction(+:total) schedule(static)
        for (int i = 0; i < ni; i++) total += c->tns.get_neighborlist(si, sj, i).size();
    }
    return total;
}

/* CSR export; `sort_lists` sorts each list ascending (what the reference comparator does,
   tests/BruteforceNSearch.cpp:135); returns the number of lists that were NOT already ascending */
int64_t tnsref_pair_export(void* h, int si, int sj, int64_t* offsets, int32_t* indices, int sort_lists)
{
    auto* c = static_cast<RefCtx*>(h);
    const int ni = c->n[si];
    offsets[0] = 0;
    if (c->last_mode == 2) {
        const auto lists = c->bf.get_neighbor_list_copy(si, sj);
        for (int i = 0; i < ni; i++) {
            offsets[i + 1] = offsets[i] + (int64_t)lists[i].size();
            std::copy(lists[i].begin(), lists[i].end(), indices + offsets[i]);
        }
        return 0;
    }
    for (int i = 0; i < ni; i++) offsets[i + 1] = offsets[i] + c->tns.get_neighborlist(si, sj, i).size();
    int64_t unsorted = 0;
    #pragma omp parallel for reduction(+:unsorted) schedule(static)
    for (int i = 0; i < ni; i++) {
        const tns::NeighborList l = c->tns.get_neighborlist(si, sj, i);
        int32_t* dst = indices + offsets[i];
        std::memcpy(dst, l.get_ptr(), sizeof(int32_t) * (size_t)l.size());
        if (!std::is_sorted(dst, dst + l.size())) {
            unsorted++;
            if (sort_lists) std::sort(dst, dst + l.size());
        }
    }
    return unsorted;
}

void tnsref_prepare_zsort(void* h) { static_cast<RefCtx*>(h)->tns.prepare_zsort(); }

void tnsref_get_zsort_order(void* h, int s, int32_t* out)
{
    auto* c = static_cast<RefCtx*>(h);
    const std::vector<int>& o = c->tns.get_zsort_order(s);
    std::memcpy(out, o.data(), sizeof(int32_t) * o.size());
}

void tnsref_apply_zsort_f32(void* h, int s, float* data, int stride) { static_cast<RefCtx*>(h)->tns.apply_zsort(s, data, stride); }

uint64_t tnsref_neighborlist_n_bytes(void* h) { return static_cast<RefCtx*>(h)->tns.get_neighborlist_n_bytes(); }

/* wall-clock milliseconds of `reps` consecutive run() calls (mode 0) or run_scalar() (mode 1), one entry per call */
void tnsref_time_runs(void* h, int mode, int reps, double* ms_out)
{
    auto* c = static_cast<RefCtx*>(h);
    for (int r = 0; r < reps; r++) {
        const auto t0 = std::chrono::steady_clock::now();
        if (mode == 0) c->tns.run(); else c->tns.run_scalar();
        const auto t1 = std::chrono::steady_clock::now();
        ms_out[r] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    c->last_mode = mode;
}

} // extern "C"
