#!/usr/bin/env python
"""Small invocations of every kernel variant for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import numpy as np  # noqa: E402
import treensearch_b200 as t  # noqa: E402
import cases  # noqa: E402

total = 0
for kernel in (0, 1):
    for build in (0, 1):
        for name in ("uniform_fixed_5000", "variable_random_sym", "three_sets_all_pairs", "clustered_blob", "duplicates", "lattice_fixed_100"):
            case = cases.GOLDEN_CASES[name]()
            eng = t.TreeNSearch()
            eng.set_option(t.TNSB_OPT_QUERY_KERNEL, kernel)
            eng.set_option(t.TNSB_OPT_BUILD, build)
            if case["radius"] is not None:
                eng.set_search_radius(case["radius"])
            keep = []
            for (p, r) in case["sets"]:
                eng.add_point_set(p, r, variable_radius=(case["radius"] is None))
                keep.append((p, r))
            for (i, j) in case["pairs"]:
                eng.set_active_search(i, j, True)
            eng.set_symmetric_search(case["symmetric"])
            eng.run()
            eng.prepare_zsort()
            eng.run()
            total += eng.stats()["n_neighbors"]
# dense blob: neighbourhoods larger than a tile / the register path, long lists
rs = np.random.RandomState(8)
pts = np.ascontiguousarray(np.concatenate([(0.5 + 0.002 * rs.standard_normal((1500, 3))), rs.random_sample((1500, 3))]).astype(np.float32))
for kernel in (0, 1):
    eng = t.TreeNSearch()
    eng.set_option(t.TNSB_OPT_QUERY_KERNEL, kernel)
    eng.set_search_radius(0.04)
    eng.add_point_set(pts)
    eng.set_active_search(0, 0, True)
    eng.run()
    total += eng.stats()["n_neighbors"]
# steady state: speculative grid reuse, graph capture + replay, host results (zero-copy, ascending lists)
case = cases.GOLDEN_CASES["uniform_fixed_5000"]()
pts5 = case["sets"][0][0].copy()
eng = t.TreeNSearch()
eng.set_search_radius(case["radius"])
eng.add_point_set(pts5)
eng.set_active_search(0, 0, True)
for _ in range(6):
    pts5 += np.float32(1e-4)
    eng.run()
total += eng.stats()["n_neighbors"] + eng.stats()["graph_replay"]
# heavy bricks with few queries (single-query warp tasks) and the latency-optimised global slow path: a sparse searching set next to a
# dense searched slab, variable radii, symmetric and not
rs = np.random.RandomState(5)
p0 = rs.random_sample((3000, 3)).astype(np.float32)
p1 = rs.random_sample((60000, 3)).astype(np.float32)
p1[:, 2] *= np.float32(0.02)
r0 = (0.05 * (1.0 + 0.5 * rs.random_sample(3000))).astype(np.float32)
r1 = (0.05 * (0.8 + 0.4 * rs.random_sample(60000))).astype(np.float32)
for sym in (True, False):
    eng = t.TreeNSearch()
    eng.add_point_set(p0, r0, variable_radius=True)
    eng.add_point_set(p1, r1, variable_radius=True)
    for (i, j) in ((0, 0), (0, 1), (1, 0)):
        eng.set_active_search(i, j, True)
    eng.set_symmetric_search(sym)
    eng.run()
    eng.run()
    total += eng.stats()["n_neighbors"] + eng.stats()["n_slow_queries"]
print("sanitize_small done, neighbours:", total)
