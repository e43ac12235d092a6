#!/usr/bin/env python
"""C4 (2M fluid + 500K boundary, variable radii) one active pair at a time: device time of every run (debugging aid, not a bench)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402
import treensearch_b200 as t  # noqa: E402
from treensearch_b200 import clouds  # noqa: E402

p0, r0, p1, r1, _ = clouds.two_set_cloud()
keep = [torch.from_numpy(x).cuda() for x in (p0, r0, p1, r1)]
for sym in (True, False):
    for pairs in (((0, 0),), ((0, 1),), ((1, 0),), ((0, 0), (0, 1), (1, 0))):
        eng = t.TreeNSearch(0)
        eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
        eng.add_point_set(keep[0], keep[1], variable_radius=True)
        eng.add_point_set(keep[2], keep[3], variable_radius=True)
        for (i, j) in pairs:
            eng.set_active_search(i, j, True)
        eng.set_symmetric_search(sym)
        ms = []
        for _ in range(4):
            eng.run()
            ms.append(round(eng.stats()["ms_query"], 3))
        st = eng.stats()
        print("sym" if sym else "asym", pairs, "ms_query per run", ms, "slow", st["n_slow_queries"], "nbrs/query", round(st["n_neighbors"] / max(st["n_queries"], 1), 1),
              "max_list", st["max_list"], flush=True)
