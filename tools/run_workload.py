#!/usr/bin/env python
"""Runs the device-resident hot path a few times on one GPU (for ncu captures; numbers printed here are never bench values)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import torch  # noqa: E402
import treensearch_b200 as t  # noqa: E402
from treensearch_b200 import clouds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--workload", default="uniform")
ap.add_argument("--zsort", action="store_true")
args = ap.parse_args()

if args.workload == "twoset":
    # C4: 2M fluid + 500K boundary, variable radii, searches 0->0, 0->1, 1->0 (BASELINE.json configs[3])
    p0, r0, p1, r1, _ = clouds.two_set_cloud()
    keep = [torch.from_numpy(x).cuda() for x in (p0, r0, p1, r1)]
    for sym in (True, False):
        eng = t.TreeNSearch(0)
        eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
        eng.add_point_set(keep[0], keep[1], variable_radius=True)
        eng.add_point_set(keep[2], keep[3], variable_radius=True)
        for (i, j) in ((0, 0), (0, 1), (1, 0)):
            eng.set_active_search(i, j, True)
        eng.set_symmetric_search(sym)
        for _ in range(args.steps):
            eng.run()
        st = eng.stats()
        print("symmetric" if sym else "asymmetric", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items() if k in ("ms_total_device", "ms_query", "n_slow_queries", "n_neighbors", "max_list", "n_reruns")})
    sys.exit(0)
if args.workload == "uniform":
    pts = clouds.uniform_cloud(args.n, 42)
    r = float(clouds.radius_for_mean_neighbors(args.n))
else:
    pts, _, r = clouds.dam_break_cloud(args.n)
    r = float(r)
d_pts = torch.from_numpy(pts).cuda()
eng = t.TreeNSearch(0)
eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
eng.set_search_radius(r)
eng.add_point_set(d_pts)
eng.set_active_search(0, 0, True)
if args.zsort:
    eng.prepare_zsort()
    eng.apply_zsort(0, d_pts, 3)
for _ in range(args.steps):
    eng.run()
st = eng.stats()
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items() if not isinstance(v, list)})
