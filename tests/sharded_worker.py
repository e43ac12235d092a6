"""Worker of the multi-GPU parity test: python -m torch.distributed.run --nproc-per-node N tests/sharded_worker.py
Every rank searches its Z-slab with the CUDA engine; lists (translated to global ids) must equal the restated oracle's
lists on the whole cloud."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from oracle import loader  # noqa: E402
from treensearch_b200 import clouds, sharded  # noqa: E402


def scenario(rank, world, local_rank, n_total, k_mean, seed, power, radius=None, exchange="auto", tiny_windows=False):
    cloud = clouds.uniform_cloud(n_total, seed)
    cloud[:, 2] = cloud[:, 2] ** power
    r = float(clouds.radius_for_mean_neighbors(n_total, k_mean)) if radius is None else radius
    per = n_total // world
    chunk = torch.from_numpy(np.ascontiguousarray(cloud[rank * per:(rank + 1) * per])).cuda()
    search = sharded.ShardedSearch(r, rank, world, local_rank, stream=torch.cuda.current_stream(), dist=dist, exchange=exchange)
    if tiny_windows:
        assert search.exchange == "p2p"
        search._open_windows(64, 64)                     # far too small: the first step must notice, grow the windows and repeat itself
    for _ in range(3):                                   # later steps reuse buffers and (balanced) cuts
        search.step(chunk, rank * per)
    assert search.n_recuts <= 2, "balanced cuts should have been reused"
    mine = search.owned_lists_global()
    port = loader.OraclePort()
    port.set_search_radius(r)
    port.add_point_set(cloud)
    port.set_active_search(0, 0, True)
    port.run(1)
    off, idx = port.csr(0, 0)
    for g, lst in mine.items():
        assert np.array_equal(lst, idx[off[g]:off[g + 1]]), f"rank {rank}: global point {g} differs"
    counts = torch.tensor([len(mine)], device="cuda")
    dist.all_reduce(counts)
    assert int(counts.item()) == n_total, "owned points do not partition the cloud"
    return len(mine), search.n_halo


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = {}
    for exchange in ("p2p", "nccl"):                     # one-sided pushes over peer memory / count + all-to-all collectives
        owned, halo = scenario(rank, world, local_rank, 200_000, 25.0, 77, 1.5, exchange=exchange)
        # slabs thinner than the halo: points are replicated to several ranks and the record buffer has to grow
        owned2, halo2 = scenario(rank, world, local_rank, 4_000, 0.0, 78, 1.0, radius=0.6 / world + 0.15, exchange=exchange)
        assert halo2 > 0
        res[exchange] = (owned, halo, owned2, halo2)
    assert res["p2p"] == res["nccl"], res
    scenario(rank, world, local_rank, 100_000, 25.0, 79, 1.0, exchange="p2p", tiny_windows=True)
    if rank == 0:
        owned, halo, owned2, halo2 = res["p2p"]
        print(f"SHARDED_OK world={world} owned={owned} halo={halo} | thin slabs: owned={owned2} halo={halo2} | p2p == nccl, window regrowth ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
