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
