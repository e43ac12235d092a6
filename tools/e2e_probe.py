#!/usr/bin/env python
"""End-to-end probe (pinned host points in, host lists out): default D2H copy vs zero-copy result writes."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402
import treensearch_b200 as t  # noqa: E402
from treensearch_b200 import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = clouds.uniform_cloud(n, 42)
r = float(clouds.radius_for_mean_neighbors(n))
h_pts = torch.from_numpy(pts).pin_memory()
for zc in (0, 1):
    eng = t.TreeNSearch(0)
    eng.set_option(t.TNSB_OPT_ZERO_COPY_RESULTS, zc)
    eng.set_search_radius(r)
    eng.add_point_set(h_pts)
    eng.set_active_search(0, 0, True)
    for it in range(5):
        t0 = time.perf_counter()
        eng.run()
        dt = (time.perf_counter() - t0) * 1e3
        st = eng.stats()
        if it >= 2:
            print(f"zero_copy={zc} wall={dt:.2f} ms upload={st['ms_upload']:.2f} device={st['ms_total_device']:.2f} query={st['ms_query']:.2f} "
                  f"download={st['ms_download']:.2f} nb={st['n_neighbors']}")
    eng.close()
