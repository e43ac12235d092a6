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
