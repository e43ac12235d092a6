#!/usr/bin/env python
"""Phase timing of one sharded step (torchrun, N ranks): where do the milliseconds around the local search go?
Every phase is bracketed by a device synchronize, so the sum is larger than a pipelined step; the point is the split."""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import torch.distributed as dist
from treensearch_b200 import sharded, clouds, _lib as L

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
job = sharded.ShardedUniformJob("uniform", int(os.environ.get("PPG", "10000000")), rank, world, lr, torch.cuda.current_stream(), os.environ.get("SHARD_INPUT", "slab"))
S = job.search
S.engine.set_option(L.TNSB_OPT_HOST_RESULTS, 0)
for _ in range(3):
    job.step_device()
import re, subprocess
def nvlink_kib(gpu):
    """Sum of the NVLink data counters (KiB transmitted, KiB received) of one GPU (nvidia-smi nvlink -gt d)."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return tx, rx
acc = {}
def tick(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0) * 1e3
    return t
K = 10
eng = S.engine
for _ in range(K):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    if S.exchange == "p2p":
        n = int(job.d_pts.shape[0])
        cuts_c = (C.c_float * (world + 1))(*[float(c) if np.isfinite(c) else 0.0 for c in S.cuts])
        parity = S._step_no & 1
        S._step_no += 1
        S._flag.fill_(0)
        eng._check(eng._lib.tnsb_shard_push(eng._h, parity, job.d_pts.data_ptr(), n, 3, int(job.id_base), S.axis, cuts_c, world, float(sharded.halo_width(S.radius)), S._flag.data_ptr()))
        t = tick("push", t)
        dist.all_reduce(S._flag, op=dist.ReduceOp.MAX); t = tick("all_reduce", t)
        ptr, n_owned, n_halo = C.c_void_p(), C.c_int64(), C.c_int64()
        eng._check(eng._lib.tnsb_shard_collect(eng._h, parity, C.byref(ptr), C.byref(n_owned), C.byref(n_halo)))
        flag = int(S._flag.item()); t = tick("collect", t)
        S.local = torch.as_tensor(sharded._DeviceRecords(ptr.value, n_owned.value + n_halo.value), device=S.device)
        S.n_owned, S.n_halo = int(n_owned.value), int(n_halo.value)
    else:
        counts = S._partition(job.d_pts, job.id_base, S.cuts, sharded.halo_width(S.radius)); t = tick("partition", t)
        S.local, S.n_owned, S.n_halo, flag = sharded.exchange_records(dist, S._records, counts, world, 0); t = tick("exchange", t)
    eng.set_option(L.TNSB_OPT_QUERY_LIMIT, S.n_owned)
    eng.resize_point_set(0, S.local, n_points=S.local.shape[0]); t = tick("resize", t)
    eng.run(); t = tick("run", t)
# NVLink bytes of K un-instrumented steps (counters of this rank's GPU) next to the algorithmic bytes of the exchange
dist.barrier(); torch.cuda.synchronize()
nv0 = nvlink_kib(lr)
for _ in range(K):
    job.step_device()
dist.barrier(); torch.cuda.synchronize()
nv1 = nvlink_kib(lr)
if rank == 0:
    if nv0 and nv1:
        tx, rx = (nv1[0] - nv0[0]) * 1024.0 / K, (nv1[1] - nv0[1]) * 1024.0 / K
        print(f"nvlink per step (rank 0 GPU): tx {tx / 1e6:.2f} MB, rx {rx / 1e6:.2f} MB; algorithmic: halo records received {S.n_halo * 16 / 1e6:.2f} MB "
              f"(+ migrating owned records), {S.n_halo} halo points of {S.n_owned} owned = {100.0 * S.n_halo / max(S.n_owned, 1):.2f} %")
    st = eng.stats()
    print(S.exchange, {k: round(v / K, 3) for k, v in acc.items()}, "n_owned", S.n_owned, "n_halo", S.n_halo,
          {k: round(st[k], 3) for k in ("ms_aabb", "ms_keys", "ms_sort", "ms_reorder", "ms_query", "ms_total_device", "ms_wall")}, "slow", st["n_slow_queries"])
dist.destroy_process_group()
