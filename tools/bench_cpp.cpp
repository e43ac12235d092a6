// bench_cpp.cpp -- end-to-end timing of the DROP-IN path: a C++ caller that holds its points in a pageable std::vector and goes
// through include/TreeNSearch exactly like a user of the reference (README.md:36-111 of the reference): run() every step, then reads
// neighbour lists on the host.  Built by treensearch_b200/build.py; bench.py runs it and reports the numbers as the `e2e` arm.
//   bench_cpp <n_points> <steps> <warmup> <pin_user_memory 0|1>
#include <TreeNSearch>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 10000000;
    const int steps = argc > 2 ? atoi(argv[2]) : 5;
    const int warmup = argc > 3 ? atoi(argv[3]) : 2;
    const int pin = argc > 4 ? atoi(argv[4]) : 0;

    // mt19937(42), uniform [0,1), xyzxyz; r for ~30 neighbours
    std::vector<float> P((size_t)3 * n);
    std::mt19937 gen(42);
    std::uniform_real_distribution<float> dist(0.0f, 1.0f);
    for (auto& v : P) v = dist(gen);
    const float r = (float)std::cbrt(30.0 / ((double)n * 4.18879020479));

    tns::TreeNSearch ns;
    if (pin) tnsb_set_option(ns.native_handle(), TNSB_OPT_PIN_USER_MEMORY, 1);
    ns.set_search_radius(r);
    const int set = ns.add_point_set(P.data(), n);
    ns.set_active_search(set, set, true);

    std::vector<double> ms;
    long long touched = 0;
    for (int k = 0; k < warmup + steps; k++) {
        // the caller moves its points between steps (in place, same pointer): every run() has to re-read them
        for (size_t i = (size_t)k; i < P.size(); i += 4099) P[i] = std::nextafter(P[i], 0.5f);
        const auto t0 = std::chrono::steady_clock::now();
        ns.run();
        // a consumer touches the lists on the host (one in 997 here: the timing is about having them addressable)
        for (int i = 0; i < n; i += 997) {
            const tns::NeighborList nl = ns.get_neighborlist(set, set, i);
            touched += nl.size() > 0 ? nl[0] : 0;
        }
        const auto t1 = std::chrono::steady_clock::now();
        if (k >= warmup) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    tnsb_stats st;
    tnsb_get_stats(ns.native_handle(), &st);
    double mean = 0.0;
    for (double v : ms) mean += v;
    mean /= (double)ms.size();
    printf("{\"n_points\": %d, \"steps\": %d, \"pin_user_memory\": %d, \"ms_mean\": %.4f, \"ms_best\": %.4f, \"n_neighbors\": %lld, "
           "\"h2d_bytes\": %lld, \"d2h_bytes\": %lld, \"ms_upload\": %.4f, \"ms_device\": %.4f, \"ms_download\": %.4f, \"checksum\": %lld}\n",
           n, steps, pin, mean, *std::min_element(ms.begin(), ms.end()), (long long)st.n_neighbors, (long long)st.h2d_bytes, (long long)st.d2h_bytes,
           st.ms_upload, st.ms_total_device, st.ms_download, touched);
    return 0;
}
