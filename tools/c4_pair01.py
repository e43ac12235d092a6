import os, sys
sys.path.insert(0, "/root/repo")
import torch
import treensearch_b200 as t
from treensearch_b200 import clouds
p0, r0, p1, r1, _ = clouds.two_set_cloud()
keep = [torch.from_numpy(x).cuda() for x in (p0, r0, p1, r1)]
eng = t.TreeNSearch(0)
eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
eng.add_point_set(keep[0], keep[1], variable_radius=True)
eng.add_point_set(keep[2], keep[3], variable_radius=True)
eng.set_active_search(0, 1, True)
eng.set_symmetric_search(True)
for _ in range(3):
    eng.run()
print(eng.stats()["ms_query"])
