#!/usr/bin/env python
"""Small-N latency probe (the reference's micro-benchmark size): wall clock and device time per run(), graph replay on / off."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402
import treensearch_b200 as t  # noqa: E402
from treensearch_b200 import clouds  # noqa: E402

pts, r = clouds.sph_lattice(9261)
for graph in ("1", "0"):
    os.environ["TNSB_GRAPH"] = graph
    for host in (True, False):
        arr = pts.copy() if host else torch.from_numpy(pts).cuda()
        eng = t.TreeNSearch(0)
        if not host:
            eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
        eng.set_search_radius(float(r))
        eng.add_point_set(arr)
        eng.set_active_search(0, 0, True)
        for _ in range(20):
            eng.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev = 0.0
        for _ in range(500):
            eng.run()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 500 * 1e3
        st = eng.stats()
        print(f"graph={graph} host_arrays={host}: wall {wall:.4f} ms/run, device {st['ms_total_device']:.4f} ms, graph_replay {st['graph_replay']}, launches {st['n_kernel_launches']}, "
              f"stages {[round(st[k], 4) for k in ('ms_aabb', 'ms_keys', 'ms_sort', 'ms_reorder', 'ms_query', 'ms_download')]}", flush=True)
        eng.close()
