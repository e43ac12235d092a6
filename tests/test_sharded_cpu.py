"""CPU test (-m "not gpu") of the multi-GPU host logic with world_size 2 over gloo: balanced cuts from the all-reduced
histogram, count bookkeeping and the one-step owned/halo record exchange of treensearch_b200/sharded.py.  The CUDA partition
kernel and the CUDA search are replaced by a numpy partition with the same semantics and by the restated oracle, so that what
is tested here is exactly the plumbing that also runs under NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from treensearch_b200 import clouds, sharded

N_TOTAL = 6000
RADIUS = 0.07


def partition_numpy(pts, id_base, axis, cuts, halo, world):
    """numpy statement of csrc/shard.cuh (slab_of / slab_count_kernel / slab_scatter_kernel)."""
    v = pts[:, axis].astype(np.float32)
    inner = cuts[1:-1].astype(np.float32)
    owner = np.searchsorted(inner, v, side="right")
    g_lo = np.searchsorted(inner, (v - np.float32(halo)).astype(np.float32), side="right")
    g_hi = np.searchsorted(inner, (v + np.float32(halo)).astype(np.float32), side="right")
    rec = np.empty((pts.shape[0], 4), dtype=np.float32)
    rec[:, :3] = pts
    rec[:, 3] = (id_base + np.arange(pts.shape[0], dtype=np.int32)).view(np.float32)
    owned = [rec[owner == g] for g in range(world)]
    halo_b = [rec[(g_lo <= g) & (g <= g_hi) & (owner != g)] for g in range(world)]
    counts = np.array([b.shape[0] for b in owned] + [b.shape[0] for b in halo_b], dtype=np.int64)
    return np.concatenate(owned + halo_b, axis=0), counts


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import loader
        cloud = clouds.uniform_cloud(N_TOTAL, 5)
        cloud[:, 2] = cloud[:, 2] ** 2                      # non-uniform along the slab axis: cuts must balance counts
        per = N_TOTAL // world
        chunk = np.ascontiguousarray(cloud[rank * per:(rank + 1) * per])
        axis, n_bins = 2, 512
        box = torch.from_numpy(np.concatenate([-chunk.min(0), chunk.max(0)]))
        dist.all_reduce(box, op=dist.ReduceOp.MAX)
        lo, hi = float(-box[axis]), float(box[3 + axis]) * (1 + 1e-6)
        hist = torch.from_numpy(np.histogram(chunk[:, axis], bins=n_bins, range=(lo, hi))[0].astype(np.int32))
        dist.all_reduce(hist)
        assert int(hist.sum()) == N_TOTAL
        cuts = sharded.balanced_cuts(hist.numpy(), lo, hi, world)
        halo = sharded.halo_width(RADIUS)
        records, counts = partition_numpy(chunk, rank * per, axis, cuts, halo, world)
        local, n_owned, n_halo, flag = sharded.exchange_records(dist, torch.from_numpy(records), counts, world, flag=rank)
        assert flag == world - 1                       # the flag that rides on the counts exchange: max over ranks
        local = local.numpy()
        # owned counts are balanced and partition the cloud
        tot = torch.tensor([n_owned, n_halo])
        gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, tot)
        owned_counts = [int(g[0]) for g in gathered]
        assert sum(owned_counts) == N_TOTAL
        assert max(owned_counts) - min(owned_counts) <= 0.05 * N_TOTAL
        # every owned point lies inside this rank's slab, every halo point within `halo` outside of it
        z = local[:, axis]
        assert np.all((z[:n_owned] >= cuts[rank]) & (z[:n_owned] < cuts[rank + 1]))
        if n_halo:
            zh = z[n_owned:]
            assert np.all((zh < cuts[rank]) | (zh >= cuts[rank + 1]))
            assert np.all((zh >= cuts[rank] - halo * 1.001) & (zh < cuts[rank + 1] + halo * 1.001))
        # local search on [owned | halo] == global search restricted to the owned points
        ids = local[:, 3].copy().view(np.int32)
        port_local = loader.OraclePort()
        port_local.set_search_radius(RADIUS)
        port_local.add_point_set(np.ascontiguousarray(local[:, :3]))
        port_local.set_active_search(0, 0, True)
        port_local.run(1)
        off, idx = port_local.csr(0, 0)
        port_glob = loader.OraclePort()
        port_glob.set_search_radius(RADIUS)
        port_glob.add_point_set(cloud)
        port_glob.set_active_search(0, 0, True)
        port_glob.run(1)
        goff, gidx = port_glob.csr(0, 0)
        for i in range(n_owned):
            mine = np.sort(ids[idx[off[i]:off[i + 1]]])
            g = int(ids[i])
            assert np.array_equal(mine, gidx[goff[g]:goff[g + 1]]), f"rank {rank}: lists of global point {g} differ"
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(n_owned))
    finally:
        dist.destroy_process_group()


def test_two_rank_slab_exchange_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_balanced_cuts_properties():
    rs = np.random.RandomState(0)
    hist = rs.randint(0, 100, 4096)
    cuts = sharded.balanced_cuts(hist, 0.0, 1.0, 8)
    assert cuts[0] == -np.inf and cuts[-1] == np.inf and np.all(np.diff(cuts[1:-1]) >= 0)
    edges = np.concatenate([[0], np.round(cuts[1:-1] * 4096).astype(int), [4096]])
    per = [hist[a:b].sum() for a, b in zip(edges[:-1], edges[1:])]
    assert max(per) - min(per) <= 2 * hist.max()
    # degenerate: everything in one bin -> all inner cuts collapse, one part owns everything
    hist = np.zeros(64, np.int64)
    hist[10] = 1000
    cuts = sharded.balanced_cuts(hist, 0.0, 64.0, 4)
    assert np.all(np.diff(cuts[1:-1]) >= 0)


def test_cost_weighted_cuts():
    """cost_weights: a uniform histogram keeps the count-balanced cuts; a density gradient moves the cuts so that the dense end gets
    fewer points and every part about the same weight."""
    uniform = np.full(4096, 250, np.int64)
    assert np.array_equal(sharded.balanced_cuts(uniform, 0.0, 1.0, 8), sharded.balanced_cuts(uniform, 0.0, 1.0, 8, sharded.cost_weights(uniform)))
    z = (np.random.RandomState(1).random_sample(1_000_000) ** 1.5)
    hist = np.histogram(z, bins=4096, range=(0.0, 1.0))[0]
    w = sharded.cost_weights(hist)
    cuts = sharded.balanced_cuts(hist, 0.0, 1.0, 8, w)
    assert cuts[0] == -np.inf and cuts[-1] == np.inf and np.all(np.diff(cuts[1:-1]) >= 0)
    edges = np.concatenate([[0], np.round(cuts[1:-1] * 4096).astype(int), [4096]])
    per_w = np.array([w[a:b].sum() for a, b in zip(edges[:-1], edges[1:])])
    per_n = np.array([hist[a:b].sum() for a, b in zip(edges[:-1], edges[1:])])
    assert per_w.max() - per_w.min() <= 2 * w.max() + 1e-9
    assert per_n[0] < per_n[-1]                      # the dense end (z -> 0) owns fewer points than the sparse end
    assert per_n.sum() == hist.sum()


def test_single_rank_exchange_is_identity():
    rec = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    counts = np.array([7, 3], dtype=np.int64)
    local, n_owned, n_halo, flag = sharded.exchange_records(None, rec, counts, 1, flag=5)
    assert n_owned == 7 and n_halo == 3 and flag == 5 and torch.equal(local, rec)
