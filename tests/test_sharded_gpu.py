"""GPU tests (-m gpu) of the multi-GPU path: the CUDA slab partition against its numpy statement, the sharded search at
world_size 1 against the oracle, and -- when the box has >= 2 GPUs -- a 2-rank NCCL run (tests/sharded_worker.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from oracle import loader
from test_sharded_cpu import partition_numpy
from treensearch_b200 import clouds, sharded

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_partition_kernel_matches_numpy(built_library):
    import torch
    import treensearch_b200 as t
    pts = clouds.uniform_cloud(50_000, 3)
    d_pts = torch.from_numpy(pts).cuda()
    eng = t.TreeNSearch(0)
    world, axis, halo = 4, 2, np.float32(0.03)
    cuts = np.array([-np.inf, 0.2, 0.45, 0.8, np.inf], dtype=np.float32)
    rec = torch.empty((80_000, 4), dtype=torch.float32, device="cuda")
    counts = (C.c_int64 * (2 * world))()
    cuts_c = (C.c_float * (world + 1))(*[float(c) if np.isfinite(c) else 0.0 for c in cuts])
    eng._check(eng._lib.tnsb_shard_partition(eng._h, d_pts.data_ptr(), pts.shape[0], 3, 1000, axis, cuts_c, world, float(halo),
                                            rec.data_ptr(), rec.shape[0], counts))
    counts = np.array(counts[:], dtype=np.int64)
    ref_rec, ref_counts = partition_numpy(pts, 1000, axis, cuts, halo, world)
    assert np.array_equal(counts, ref_counts)
    got = rec[: counts.sum()].cpu().numpy()
    start = 0
    for b in range(2 * world):                     # same records per bucket, order inside a bucket is free
        a = got[start:start + counts[b]]
        e = ref_rec[start:start + counts[b]]
        ka = np.argsort(a[:, 3].copy().view(np.int32))
        ke = np.argsort(e[:, 3].copy().view(np.int32))
        assert np.array_equal(a[ka].view(np.int32), e[ke].view(np.int32)), f"bucket {b}"
        start += counts[b]
    # aabb + histogram helpers
    mm = (C.c_float * 6)()
    eng._check(eng._lib.tnsb_shard_aabb(eng._h, d_pts.data_ptr(), pts.shape[0], 3, mm))
    assert np.array_equal(np.array(mm[:3], np.float32), pts.min(0)) and np.array_equal(np.array(mm[3:], np.float32), pts.max(0))
    hist = torch.zeros(256, dtype=torch.int32, device="cuda")
    eng._check(eng._lib.tnsb_shard_histogram(eng._h, d_pts.data_ptr(), pts.shape[0], 3, axis, 0.0, 1.0, 256, hist.data_ptr()))
    torch.cuda.synchronize()
    ref_hist = np.bincount(np.clip(((pts[:, axis] - np.float32(0.0)) * np.float32(256.0)).astype(np.int32), 0, 255), minlength=256)
    assert np.array_equal(hist.cpu().numpy(), ref_hist)


def test_sharded_world1_matches_oracle(built_library):
    import torch
    pts = clouds.uniform_cloud(30_000, 9)
    r = float(clouds.radius_for_mean_neighbors(30_000))
    s = sharded.ShardedSearch(r, 0, 1, 0, stream=torch.cuda.current_stream())
    s.step(torch.from_numpy(pts).cuda(), 0)
    assert s.n_owned == 30_000 and s.n_halo == 0
    mine = s.owned_lists_global()
    case = dict(sets=[(pts, None)], radius=r, pairs=[(0, 0)], symmetric=True)
    port = cases.configure(loader.OraclePort(), case)
    port.run(1)
    off, idx = port.csr(0, 0)
    for g in range(30_000):
        assert np.array_equal(mine[g], idx[off[g]:off[g + 1]])


def test_two_rank_nccl(built_library):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "sharded_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0 and "SHARDED_OK world=2" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]


# ---------------------------------------------------------------------------------------------------- emulated ranks on ONE GPU
def _emulated_rank_lists(t, pts, r, world, axis, cuts, chunks):
    """Runs the slab partition kernel on every chunk, concatenates [owned | halo] per emulated rank exactly like the exchange step
    would deliver them, searches every rank's records with the engine (halo points find-only) and returns
    {global id: sorted global neighbour ids} of all owned points."""
    import torch
    halo = float(sharded.halo_width(r))
    eng = t.TreeNSearch(0)
    cuts_c = (C.c_float * (world + 1))(*[float(c) if np.isfinite(c) else 0.0 for c in cuts])
    owned = [[] for _ in range(world)]
    halos = [[] for _ in range(world)]
    base = 0
    for chunk in chunks:
        d = torch.from_numpy(np.ascontiguousarray(chunk)).cuda()
        rec = torch.empty((int(chunk.shape[0] * (1 + world)) + 1024, 4), dtype=torch.float32, device="cuda")
        counts = (C.c_int64 * (2 * world))()
        eng._check(eng._lib.tnsb_shard_partition(eng._h, d.data_ptr(), chunk.shape[0], 3, base, axis, cuts_c, world, halo,
                                                rec.data_ptr(), rec.shape[0], counts))
        counts = np.array(counts[:], dtype=np.int64)
        start = 0
        for b in range(2 * world):
            part = rec[start:start + counts[b]]
            (owned if b < world else halos)[b % world].append(part)
            start += counts[b]
        base += chunk.shape[0]
    out = {}
    n_owned_total = 0
    n_halo_total = 0
    for g in range(world):
        own = torch.cat(owned[g]) if owned[g] else torch.empty((0, 4), dtype=torch.float32, device="cuda")
        hal = torch.cat(halos[g]) if halos[g] else torch.empty((0, 4), dtype=torch.float32, device="cuda")
        local = torch.cat([own, hal]).contiguous()
        n_owned, n_halo = own.shape[0], hal.shape[0]
        n_owned_total += n_owned
        n_halo_total += n_halo
        e = t.TreeNSearch(0)
        e.set_search_radius(r)
        e.set_option(t.TNSB_OPT_POINT_STRIDE, 4)
        e.set_option(t.TNSB_OPT_QUERY_LIMIT, n_owned)
        e._query_limit = n_owned
        e.add_point_set(local, n_points=local.shape[0])
        e.set_active_search(0, 0, True)
        e.run()
        ids = local[:, 3].contiguous().view(torch.int32).cpu().numpy()
        ragged, pos = e.neighbor_lists(0, 0)
        for i in range(n_owned):
            p = int(pos[i])
            n = int(ragged[p])
            assert int(ids[i]) not in out, "a point is owned by two ranks"
            out[int(ids[i])] = np.sort(ids[ragged[p + 1: p + 1 + n]])
    return out, n_owned_total, n_halo_total


@pytest.mark.parametrize("cloud", ["uniform", "clustered_z", "thin_slabs"])
def test_emulated_ranks_match_reference(built_library, cloud):
    """The multi-GPU data path -- slab partition, [owned | halo] records with global ids, find-only halo points -- on ONE GPU:
    4 (or 16) emulated ranks, every chunk partitioned by the CUDA kernel, global neighbour sets against the unmodified reference."""
    import treensearch_b200 as t
    n = 200_000
    rs = np.random.RandomState(31)
    pts = clouds.uniform_cloud(n, 31).copy()
    world = 4
    if cloud == "clustered_z":
        pts[:, 2] = pts[:, 2] ** np.float32(1.5)
    if cloud == "thin_slabs":
        world = 16                                    # slabs thinner than the halo: points replicated to several ranks
        pts[:, 2] *= np.float32(0.25)
    r = float(clouds.radius_for_mean_neighbors(n, 30.0, volume=0.25 if cloud == "thin_slabs" else 1.0))
    axis = 2
    lo, hi = float(pts[:, axis].min()), float(pts[:, axis].max())
    hi += max(1e-6 * abs(hi - lo), 1e-30)
    hist = np.bincount(np.clip(((pts[:, axis] - np.float32(lo)) * np.float32(4096.0 / (hi - lo))).astype(np.int64), 0, 4095), minlength=4096)
    cuts = sharded.balanced_cuts(hist, lo, hi, world)
    perm = rs.permutation(n)                          # every "rank" starts with an arbitrary chunk of the cloud
    chunks = [pts[perm[k::4]] for k in range(4)]
    shuffled = np.concatenate(chunks)                 # global id = position in this concatenation
    mine, n_owned, n_halo = _emulated_rank_lists(t, pts, r, world, axis, cuts, chunks)
    assert n_owned == n and n_halo > 0
    case = dict(sets=[(np.ascontiguousarray(shuffled), None)], radius=r, pairs=[(0, 0)], symmetric=True)
    ref = cases.configure(loader.Reference() if loader.reference_available() else loader.OraclePort(), case)
    ref.run(0 if loader.reference_available() else 1)
    off, idx = ref.csr(0, 0)
    assert len(mine) == n
    for g in range(n):
        assert np.array_equal(mine[g], idx[off[g]:off[g + 1]]), f"global point {g}"


def test_sharded_default_stream(built_library):
    """ShardedSearch without an explicit stream runs on torch's current stream (the exchange must be stream ordered)."""
    import torch
    pts = clouds.uniform_cloud(20_000, 5)
    r = float(clouds.radius_for_mean_neighbors(20_000))
    s = sharded.ShardedSearch(r, 0, 1, 0)
    assert s.stream.cuda_stream == torch.cuda.current_stream().cuda_stream
    s.step(torch.from_numpy(pts).cuda(), 0)
    assert s.n_owned == 20_000
