#!/usr/bin/env python
"""Phase timing of one sharded step (torchrun, N ranks): where do the milliseconds around the local search go?"""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import torch.distributed as dist
from treensearch_b200 import sharded, clouds, _lib as L

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
job = sharded.ShardedUniformJob("uniform", 10_000_000, rank, world, lr, torch.cuda.current_stream(), "slab")
S = job.search
S.engine.set_option(L.TNSB_OPT_HOST_RESULTS, 0)
for _ in range(3):
    job.step_device()
acc = {}
def tick(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0) * 1e3
    return t
K = 10
for _ in range(K):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    counts = S._partition(job.d_pts, job.id_base, S.cuts, sharded.halo_width(S.radius)); t = tick("partition", t)
    S.local, S.n_owned, S.n_halo, flag = sharded.exchange_records(dist, S._records, counts, world, 0); t = tick("exchange", t)
    S.engine.set_option(L.TNSB_OPT_QUERY_LIMIT, S.n_owned)
    S.engine.resize_point_set(0, S.local, n_points=S.local.shape[0]); t = tick("resize", t)
    S.engine.run(); t = tick("run", t)
if rank == 0:
    print({k: round(v / K, 3) for k, v in acc.items()}, "n_owned", S.n_owned, "n_halo", S.n_halo, S.engine.stats()["ms_total_device"])
dist.destroy_process_group()
