#pragma once
// tns::TreeNSearch -- the reference's public class (TreeNSearch/source/TreeNSearch.h:28-335) re-implemented as a thin,
// header-only forward to the C ABI of the B200 engine (include/tnsb.h, libtnsb.so).  Signatures are kept verbatim so that
// callers such as SPH solvers and the reference's own tests (tests/tests.cpp, tests/BruteforceNSearch.cpp) compile
// unchanged:   g++ ... -I<repo>/include  -L<repo>/treensearch_b200 -ltnsb
//
// Error behaviour follows the reference: a message on std::cout, then exit(-1) (TreeNSearch.cpp:22-25, :366-392).
// Unlike the reference header this one leaks neither <immintrin.h> types nor Taskflow.
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../tnsb.h"
#include "NeighborList.h"

namespace tns
{
	class TreeNSearch
	{
	public:
		// -----------------------------------------------  CONSTRUCTORS  -----------------------------------------------
		TreeNSearch()
		{
			if (tnsb_create(&ctx_, -1) != TNSB_OK) {
				std::cout << tnsb_last_error(nullptr) << std::endl;
				exit(-1);
			}
		}
		~TreeNSearch() { tnsb_destroy(ctx_); }
		TreeNSearch(const TreeNSearch&) = delete;
		TreeNSearch& operator=(const TreeNSearch&) = delete;

		// -----------------------------------------------  MAIN INTERFACE  -----------------------------------------------
		int add_point_set(const float* points_begin, const int n_points) { return added(tnsb_add_point_set_f32(ctx_, points_begin, nullptr, n_points, 0)); }
		int add_point_set(const double* points_begin, const int n_points) { return added(tnsb_add_point_set_f64(ctx_, points_begin, nullptr, n_points, 0)); }
		void resize_point_set(const int set_id, const float* points_begin, const int n_points) { ok(tnsb_resize_point_set_f32(ctx_, set_id, points_begin, nullptr, n_points, 0)); invalidate(); }
		void resize_point_set(const int set_id, const double* points_begin, const int n_points) { ok(tnsb_resize_point_set_f64(ctx_, set_id, points_begin, nullptr, n_points, 0)); invalidate(); }
		void set_search_radius(const float search_radius) { ok(tnsb_set_search_radius(ctx_, search_radius)); }
		void set_search_radius(const double search_radius) { this->set_search_radius((float)search_radius); }
		int add_point_set(const float* points_begin, const float* radii_begin, const int n_points) { return added(tnsb_add_point_set_f32(ctx_, points_begin, radii_begin, n_points, 1)); }
		int add_point_set(const double* points_begin, const double* radii_begin, const int n_points) { return added(tnsb_add_point_set_f64(ctx_, points_begin, radii_begin, n_points, 1)); }
		void resize_point_set(const int set_id, const float* points_begin, const float* radii_begin, const int n_points) { ok(tnsb_resize_point_set_f32(ctx_, set_id, points_begin, radii_begin, n_points, 1)); invalidate(); }
		void resize_point_set(const int set_id, const double* points_begin, const double* radii_begin, const int n_points) { ok(tnsb_resize_point_set_f64(ctx_, set_id, points_begin, radii_begin, n_points, 1)); invalidate(); }
		void set_cell_size(const float cell_size) { ok(tnsb_set_cell_size(ctx_, cell_size)); }
		void set_cell_size(const double cell_size) { this->set_cell_size((float)cell_size); }

		/** Runs the whole search on the GPU and mirrors the neighbour lists into host memory. */
		void run()
		{
			if (n_threads_ == -1) { n_threads_ = max_threads(); }   // TreeNSearch.cpp:266-268
			if (recursion_cap_ <= 0) { fail("TreeNSearch error: n_points_to_stop_recursion <= 0."); }   // TreeNSearch.cpp:372-375
			ok(tnsb_run(ctx_));
			fetch_views();
		}

		inline NeighborList get_neighborlist(const int set_i, const int set_j, const int point_i) const
		{
			const PairView& v = views_[(size_t)set_i * (size_t)n_sets_ + (size_t)set_j];
			return NeighborList(v.list_pos32 ? v.ragged + v.list_pos32[point_i] : v.ragged + v.list_pos[point_i]);
		}

		template<typename FUNC>
		inline void for_each_neighbor(const int set_i, const int set_j, const int i, FUNC f)
		{
			const NeighborList neighbors = this->get_neighborlist(set_i, set_j, i);
			const int n = neighbors.size();
			for (int loc_j = 0; loc_j < n; loc_j++) { f(neighbors[loc_j]); }
		}

		void prepare_zsort()
		{
			if (n_threads_ == -1) { n_threads_ = max_threads(); }
			ok(tnsb_prepare_zsort(ctx_));
			const int n_sets = this->get_n_sets();
			zsort_.resize((size_t)n_sets);
			for (int s = 0; s < n_sets; s++) {
				const int32_t* order = nullptr; int n = 0;
				ok(tnsb_get_zsort_order(ctx_, s, &order, &n));
				zsort_[(size_t)s].assign(order, order + n);
			}
		}

		/** In-place gather data[new] = data[old] of any per-point array (same contract as TreeNSearch.h:443-481). */
		template<typename T>
		void apply_zsort(const int set_i, T* data_ptr, const int stride = 1) const
		{
			if (!this->does_set_exist(set_i)) { fail("tns::TreeNSearch::apply_zsort error: set to z_sort does not exit."); }
			if ((size_t)set_i >= zsort_.size()) {
				std::cout << "tns::TreeNSearch::apply_zsort error: no zsort order ready for set_i (" << set_i << ")." << std::endl;
				exit(-1);
			}
			const std::vector<int>& new_to_old = zsort_[(size_t)set_i];
			const long long n_points = (long long)this->get_n_points_in_set(set_i);
			const long long s = stride;
			std::vector<T> tmp((size_t)(n_points * s));
			#pragma omp parallel for schedule(static) num_threads(this->threads())
			for (long long i = 0; i < n_points * s; i++) { tmp[(size_t)i] = data_ptr[i]; }
			#pragma omp parallel for schedule(static) num_threads(this->threads())
			for (long long new_idx = 0; new_idx < n_points; new_idx++) {
				const long long old_idx = new_to_old[(size_t)new_idx];
				for (long long j = 0; j < s; j++) { data_ptr[new_idx * s + j] = tmp[(size_t)(old_idx * s + j)]; }
			}
		}

		void set_symmetric_search(const bool activate) { ok(tnsb_set_symmetric_search(ctx_, activate ? 1 : 0)); }

		// -----------------------------------------------  SECONDARY METHODS  -----------------------------------------------
		/** The reference's scalar twin of run() (TreeNSearch.cpp:150-160).  Here: the same CUDA path. */
		void run_scalar() { this->run(); }

		void print_state() const
		{
			tnsb_stats st;
			tnsb_get_stats(ctx_, &st);
			std::cout << "\n ================ OPTIONS ================ " << std::endl;
			std::cout << "n_points_to_stop_recursion: " << recursion_cap_ << " (unused: no octree)" << std::endl;
			std::cout << "n_threads: " << n_threads_ << std::endl;
			std::cout << "\n ================ GRID ================ " << std::endl;
			std::cout << "World AABB float" << std::endl;
			std::cout << "[" << st.domain_bottom[0] << ", " << st.domain_bottom[1] << ", " << st.domain_bottom[2] << "]" << std::endl;
			std::cout << "[" << st.domain_top[0] << ", " << st.domain_top[1] << ", " << st.domain_top[2] << "]" << std::endl;
			std::cout << "cell_size: " << st.cell_size << std::endl;
			std::cout << "# cells: " << st.n_cells << std::endl;
			std::cout << "morton key bits: " << st.key_bits << ", radix sort passes: " << st.sort_passes << std::endl;
			std::cout << "device time (ms): " << st.ms_total_device << " (query " << st.ms_query << ", sort " << st.ms_sort << ")" << std::endl;
			std::cout << "\n ================ NEIGHBORLISTS ================ " << std::endl;
			std::cout << "Active searches: " << std::endl;
			const int n_sets = this->get_n_sets();
			for (int i = 0; i < n_sets; i++)
				for (int j = 0; j < n_sets; j++)
					if (this->is_search_active(i, j)) { std::cout << "\t" << "set_" << i << " -> " << "set_" << j << std::endl; }
			std::cout << "Total memory (MB): " << (double)this->get_neighborlist_n_bytes() / 1024.0 / 1024.0 << std::endl;
			std::cout << "\n ================ PER SET DATA ================ " << std::endl;
			for (int i = 0; i < n_sets; i++) {
				std::cout << "\n ---------------- set_" << i << " ---------------- " << std::endl;
				std::cout << "# points: " << this->get_n_points_in_set(i) << std::endl;
				for (int j = 0; j < n_sets; j++) {
					int64_t mms[3];
					if (this->is_search_active(i, j) && tnsb_get_pair_neighbor_stats(ctx_, i, j, mms) == TNSB_OK) {
						const int n = this->get_n_points_in_set(i);
						std::cout << "n_neighbors set_" << i << " -> " << "set_" << j << " [min, max, avg]: [" << mms[0] << ", " << mms[1] << ", "
							<< (n > 0 ? (double)mms[2] / (double)n : 0.0) << "]" << std::endl;
					}
				}
			}
		}

		uint64_t get_neighborlist_n_bytes() const { return tnsb_get_neighborlist_n_bytes(ctx_); }

		// -----------------------------------------------  SETTERS AND GETTERS  -----------------------------------------------
		void set_all_searches(const bool active) { ok(tnsb_set_all_searches(ctx_, active ? 1 : 0)); }
		void set_active_search(const int set_i, const int set_j, const bool active = true) { ok(tnsb_set_active_search(ctx_, set_i, set_j, active ? 1 : 0)); }
		void set_active_search(const int set_i, const bool search_in_all = true, const bool be_found_by_all = true) { ok(tnsb_set_active_search_of_set(ctx_, set_i, search_in_all ? 1 : 0, be_found_by_all ? 1 : 0)); }

		void set_n_threads(const int n_threads) { n_threads_ = n_threads; }                    // host-side loops only (apply_zsort)
		void set_recursion_cap(const int cap) { recursion_cap_ = cap; }                        // accepted for compatibility: no octree here
		void set_n_points_for_parallel_octree(const int n_points = 200000) { (void)n_points; } // accepted for compatibility

		int get_n_sets() const { return tnsb_get_n_sets(ctx_); }
		int get_n_threads() const { return n_threads_; }
		int get_n_points_in_set(const int set_i) const { return tnsb_get_n_points_in_set(ctx_, set_i); }
		int get_total_n_points() const { return tnsb_get_total_n_points(ctx_); }
		bool is_search_active(const int set_i, const int set_j) const { return tnsb_is_search_active(ctx_, set_i, set_j) != 0; }
		bool does_set_exist(const int set_i) const { return tnsb_does_set_exist(ctx_, set_i) != 0; }
		const std::vector<int>& get_zsort_order(const int set_i) const { return zsort_[(size_t)set_i]; }

		// -----------------------------------------------  ENGINE EXTRAS (not in the reference)  -----------------------------------------------
		/** Raw C-ABI handle, e.g. for tnsb_set_option / tnsb_get_stats / the device-resident accessors. */
		tnsb_context* native_handle() const { return ctx_; }

	private:
		struct PairView { const int* ragged = nullptr; const int64_t* list_pos = nullptr; const uint32_t* list_pos32 = nullptr; };

		[[noreturn]] static void fail(const char* msg) { std::cout << msg << std::endl; exit(-1); }
		void ok(const int rc) const { if (rc < 0) { fail(tnsb_last_error(ctx_)); } }
		int added(const int rc) { ok(rc); invalidate(); return rc; }
		void invalidate() { views_.clear(); n_sets_ = 0; }
		static int max_threads()
		{
#ifdef _OPENMP
			return omp_get_max_threads();
#else
			return 1;
#endif
		}
		int threads() const { return n_threads_ > 0 ? n_threads_ : 1; }
		void fetch_views()
		{
			n_sets_ = this->get_n_sets();
			views_.assign((size_t)n_sets_ * (size_t)n_sets_, PairView());
			for (int i = 0; i < n_sets_; i++) {
				for (int j = 0; j < n_sets_; j++) {
					if (!this->is_search_active(i, j)) { continue; }
					PairView v; int64_t n_ints = 0;
					// 32-bit positions when the engine mirrored them that way (half the PCIe bytes), else the 64-bit array
					if (tnsb_get_neighborlists_u32(ctx_, i, j, &v.ragged, &v.list_pos32, &n_ints) != TNSB_OK) {
						v.list_pos32 = nullptr;
						ok(tnsb_get_neighborlists(ctx_, i, j, &v.ragged, &v.list_pos, &n_ints));
					}
					views_[(size_t)i * (size_t)n_sets_ + (size_t)j] = v;
				}
			}
		}

		tnsb_context* ctx_ = nullptr;
		std::vector<PairView> views_;
		std::vector<std::vector<int>> zsort_;
		int n_sets_ = 0;
		int n_threads_ = -1;
		int recursion_cap_ = 1000;
	};
}
