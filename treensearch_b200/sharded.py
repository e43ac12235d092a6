"""Multi-GPU path: Z-slab sharding of one point cloud across the GPUs of a box, one process per GPU.

The reference is a single shared-memory process; this decomposition is new design.  Per step and per rank:

  1. world box of the local chunk (CUDA reduction)                       -> all_reduce(MAX) of [-min, max]      (24 bytes)
  2. histogram of the slab axis over the world extent (CUDA kernel)      -> all_reduce(SUM)                     (16 KB)
     -> the same balanced cuts on every rank (equal point counts per slab)
  3. bucket partition of the local chunk into (x, y, z, global id) records, ordered
        [owned by rank 0 | ... | owned by rank G-1 | halo of rank 0 | ... | halo of rank G-1]   (two CUDA kernels)
  4. ONE exchange step: all_to_all of the per-destination counts, then all_to_all_single of the owned records and of
     the halo records (NCCL over NVLink / NVSwitch).  A halo of width >= r_max on both sides is all a fixed-radius
     query needs, so there is no second exchange.
  5. the single-GPU engine on [owned | halo] records with TNSB_OPT_QUERY_LIMIT = n_owned: halo points are find-only.

Steps 3 + 4 exist in two forms (`exchange=`):
  "nccl"  two partition kernels, all_to_all of the counts, all_to_all_single of the owned and of the halo records;
  "p2p"   (default on CUDA with world > 1) ONE kernel pushes every record straight into its owner's receive window -- and into
          the windows that need it as halo -- over NVLink peer memory (CUDA IPC, tnsb_shard_window_* / tnsb_shard_push), followed
          by a 4-byte all_reduce that is both the barrier and the carrier of the re-balance flag.  No count pass, no count
          exchange, no all-to-all; the engine then searches the records in place in the window.

Neighbour lists come back in LOCAL indices (into the rank's [owned | halo] array, which is what a distributed consumer
indexes anyway); `local_ids` maps them to global point ids.  No data-path collective other than step 4 exists.

The host logic (cuts, count bookkeeping, exchange) is backend agnostic and is exercised on CPU with gloo in
tests/test_sharded_cpu.py; the partition / search themselves need the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def cost_weights(hist: np.ndarray) -> np.ndarray:
    """Work estimate per histogram bin for the cut placement: a bin of c points along the slab axis (uniform cross-section) has density
    ~ c, and a point's search cost grows with the density around it, so the work of the bin grows like c * (1 + c / mean) -- points
    where the cloud is as dense as on average, pairs where it is much denser.  A uniform cloud gets the count-balanced cuts."""
    h = np.asarray(hist, dtype=np.float64)
    occupied = h[h > 0]
    mean = float(occupied.mean()) if occupied.size else 1.0
    return h * (1.0 + h / mean) * 0.5


def balanced_cuts(hist: np.ndarray, lo: float, hi: float, n_parts: int, weights: np.ndarray | None = None) -> np.ndarray:
    """Cut coordinates (float32, length n_parts + 1) such that every part holds ~ the same number of points -- or, with `weights`
    (one value per bin, e.g. cost_weights(hist)), the same share of the weight.
    hist: global histogram over n_bins equal bins of [lo, hi).  cuts[0] = -inf, cuts[-1] = +inf; inner cuts lie on bin edges."""
    hist = np.asarray(hist, dtype=np.int64) if weights is None else np.asarray(weights, dtype=np.float64)
    n_bins = hist.shape[0]
    total = hist.sum()
    cum = np.cumsum(hist)
    cuts = np.empty(n_parts + 1, dtype=np.float32)
    cuts[0], cuts[-1] = -np.inf, np.inf
    width = (float(hi) - float(lo)) / n_bins
    prev = 0
    for g in range(1, n_parts):
        target = (total * g) // n_parts if weights is None else total * g / n_parts
        b = int(np.searchsorted(cum, target, side="left")) + 1      # cut after the bin that reaches the target
        b = min(max(b, prev), n_bins)
        prev = b
        cuts[g] = np.float32(float(lo) + b * width)
    return cuts


def halo_width(r_max: float) -> np.float32:
    """Halo strictly wider than any search distance (float rounding of d2 / r^2 included)."""
    return np.float32(np.float32(r_max) * np.float32(1.0 + 1.0 / 1024.0))


def exchange_records(dist, records, counts: np.ndarray, world: int, flag: int = 0):
    """Step 4.  `records` is a torch tensor [total, 4] laid out as the partition produced it; counts = int64[2*world]
    (owned counts then halo counts, per destination).  `flag` is one extra integer every rank tells every other rank in the
    same (small) counts exchange; the maximum over ranks comes back.
    Returns (local [n_owned + n_halo, 4], n_owned, n_halo, max_flag)."""
    import torch
    dev = records.device
    sc = np.empty((world, 3), dtype=np.int64)                 # [dest, (owned, halo, flag)]
    sc[:, 0], sc[:, 1], sc[:, 2] = counts[:world], counts[world:], flag
    send_counts = torch.from_numpy(sc).to(dev)
    recv_counts = torch.empty_like(send_counts)
    if world > 1:
        dist.all_to_all_single(recv_counts, send_counts)
    else:
        recv_counts.copy_(send_counts)
    rc = recv_counts.cpu().numpy()                      # [source, (owned, halo, flag)]
    own_in, halo_in = rc[:, 0].astype(np.int64), rc[:, 1].astype(np.int64)
    own_out, halo_out = counts[:world].astype(np.int64), counts[world:].astype(np.int64)
    n_owned, n_halo = int(own_in.sum()), int(halo_in.sum())
    local = torch.empty((n_owned + n_halo, 4), dtype=records.dtype, device=dev)
    owned_src = records[: int(own_out.sum())]
    halo_src = records[int(own_out.sum()): int(own_out.sum() + halo_out.sum())]
    if world > 1:
        dist.all_to_all_single(local[:n_owned], owned_src, output_split_sizes=own_in.tolist(), input_split_sizes=own_out.tolist())
        dist.all_to_all_single(local[n_owned:], halo_src, output_split_sizes=halo_in.tolist(), input_split_sizes=halo_out.tolist())
    else:
        local[:n_owned].copy_(owned_src)
        local[n_owned:].copy_(halo_src)
    return local, n_owned, n_halo, int(rc[:, 2].max())


class _DeviceRecords:
    """(n, 4) float32 view of device memory owned by the engine (a receive window), exposed through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n), 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


class ShardedSearch:
    """One rank of the Z-slab sharded fixed-radius search.  `dist` is torch.distributed (already initialised) or None."""

    def __init__(self, radius: float, rank: int = 0, world: int = 1, device: int = 0, axis: int = 2, n_bins: int = 4096,
                 stream=None, dist=None, exchange: str = "auto"):
        import torch
        from .api import TreeNSearch
        self.torch = torch
        self.dist = dist
        self.rank, self.world, self.axis, self.n_bins = rank, world, axis, n_bins
        self.radius = float(np.float32(radius))
        self.device = torch.device("cuda", device)
        self.engine = TreeNSearch(device)
        # The exchange step orders the push kernel, the barrier all_reduce and the collect on ONE stream: torch's current stream
        # unless the caller names another one (the engine's private non-blocking stream would not be ordered behind torch's NCCL calls).
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self.stream = stream
        self.engine.set_stream(stream.cuda_stream)
        self.engine.set_search_radius(self.radius)
        self.engine.set_option(L.TNSB_OPT_POINT_STRIDE, 4)
        self._set_added = False
        self._hist = torch.zeros(n_bins, dtype=torch.int32, device=self.device)
        self._records = None
        self.local = None
        self.n_owned = self.n_halo = 0
        self.cuts = None
        # temporal coherence: the cuts of the previous step are kept while every rank's owned count stays within
        # `rebalance_tolerance` of the mean; any rank can request new cuts through the flag that rides on the counts exchange
        self.rebalance_tolerance = 0.10
        self.balance = "cost"          # cut placement: "cost" (cost_weights: dense slabs get fewer points) or "count"
        self._expected_owned = None    # owned points of this rank under the current cuts (from the histogram they were made from)
        self.n_global = 0
        self._recut = True
        self.n_recuts = 0
        # one-sided exchange over peer memory needs one process per GPU talking NCCL; everything else uses the collective form
        if exchange == "auto":
            import os
            exchange = os.environ.get("TNSB_SHARD_EXCHANGE", "auto")      # "nccl" forces the collective form (e.g. ranks on several nodes)
        if exchange == "auto":
            exchange = "p2p" if (dist is not None and world > 1 and dist.get_backend() == "nccl") else "nccl"
        self.exchange = exchange
        self._win_caps = None
        self._step_no = 0
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    # ---- one-sided exchange (tnsb_shard_window_*)
    def _open_windows(self, cap_owned: int, cap_halo: int):
        """Same capacities on every rank (the pushers check them); collective."""
        torch, dist = self.torch, self.dist
        caps = torch.tensor([int(cap_owned), int(cap_halo)], dtype=torch.int64, device=self.device)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX)          # also: every rank is done with the old windows
        cap_owned, cap_halo = (int(v) for v in caps.cpu().tolist())
        mine = (C.c_ubyte * 128)()
        self.engine._check(self.engine._lib.tnsb_shard_window_create(self.engine._h, cap_owned, cap_halo, mine))
        h = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=self.device)
        allh = torch.empty(self.world * 128, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, h)
        buf = (C.c_ubyte * (self.world * 128))(*allh.cpu().tolist())
        self.engine._check(self.engine._lib.tnsb_shard_window_open(self.engine._h, self.world, self.rank, buf))
        dist.barrier()                       # nobody pushes before every rank has mapped every window
        self._win_caps = (cap_owned, cap_halo)
        self._step_no = 0

    def _push_exchange(self, points, id_base, want):
        """Partition + exchange as one kernel over peer memory.  Returns (local records view, n_owned, n_halo, flag)."""
        torch, dist = self.torch, self.dist
        n = int(points.shape[0])
        if self._win_caps is None:
            self._open_windows(int(n * 1.5) + 4096, int(n * 0.5) + 4096)
        cuts_c = (C.c_float * (self.world + 1))(*[float(c) if np.isfinite(c) else 0.0 for c in self.cuts])
        for _attempt in range(4):
            parity = self._step_no & 1
            self._step_no += 1
            self._flag.fill_(int(want))
            self.engine._check(self.engine._lib.tnsb_shard_push(self.engine._h, parity, points.data_ptr(), n, int(points.shape[1]), int(id_base), self.axis,
                                                              cuts_c, self.world, float(halo_width(self.radius)), self._flag.data_ptr()))
            dist.all_reduce(self._flag, op=dist.ReduceOp.MAX)       # barrier (stream ordered behind the push) + re-balance / overflow flag
            ptr, n_owned, n_halo, flag_c = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int(0)
            rc = self.engine._lib.tnsb_shard_collect_flag(self.engine._h, parity, C.byref(ptr), C.byref(n_owned), C.byref(n_halo),
                                                          self._flag.data_ptr(), C.byref(flag_c))      # counts + flag: ONE host round trip
            flag = int(flag_c.value)
            if flag < 2:
                self.engine._check(rc)
                local = torch.as_tensor(_DeviceRecords(ptr.value, n_owned.value + n_halo.value), device=self.device)
                return local, int(n_owned.value), int(n_halo.value), flag
            # a window overflowed somewhere: EVERY rank repeats the step with larger windows (the counts are exact)
            if rc not in (L.TNSB_OK, L.TNSB_ERR_LIMIT):
                self.engine._check(rc)
            self._open_windows(int(n_owned.value * 1.25) + 4096, int(n_halo.value * 1.5) + 4096)
        raise RuntimeError("sharded search: receive windows kept overflowing")

    # ---- thin wrappers of the shard helpers of the C ABI
    def _aabb(self, pts):
        out = (C.c_float * 6)()
        self.engine._check(self.engine._lib.tnsb_shard_aabb(self.engine._h, pts.data_ptr(), pts.shape[0], pts.shape[1], out))
        return np.array(out[:], dtype=np.float32)

    def _histogram(self, pts, lo, hi):
        self.engine._check(self.engine._lib.tnsb_shard_histogram(self.engine._h, pts.data_ptr(), pts.shape[0], pts.shape[1], self.axis,
                                                                float(lo), float(hi), self.n_bins, self._hist.data_ptr()))

    def _partition(self, pts, id_base, cuts, halo):
        torch = self.torch
        need = int(pts.shape[0] * 1.3) + 1024
        if self._records is None or self._records.shape[0] < need:
            self._records = torch.empty((need, 4), dtype=torch.float32, device=self.device)
        counts = (C.c_int64 * (2 * self.world))()
        cuts_c = (C.c_float * (self.world + 1))(*[float(c) if np.isfinite(c) else 0.0 for c in cuts])
        for _attempt in range(2):
            rc = self.engine._lib.tnsb_shard_partition(self.engine._h, pts.data_ptr(), pts.shape[0], pts.shape[1], int(id_base), self.axis,
                                                      cuts_c, self.world, float(halo), self._records.data_ptr(), self._records.shape[0], counts)
            if rc != L.TNSB_ERR_LIMIT:
                break
            # thin slabs (halo wider than a slab) replicate points to several ranks: the counts are exact, grow and retry
            need = int(sum(counts[:])) + 1024
            self._records = torch.empty((need, 4), dtype=torch.float32, device=self.device)
        self.engine._check(rc)
        return np.array(counts[:], dtype=np.int64)

    def step(self, points, id_base: int):
        """points: float32 CUDA tensor [n_local, 3] (this rank's chunk of the global cloud, ids id_base .. id_base + n_local)."""
        torch, dist = self.torch, self.dist
        if self._recut or self.cuts is None:
            # 1. world box
            mm = self._aabb(points)
            box = torch.from_numpy(np.concatenate([-mm[:3], mm[3:]])).to(self.device)
            if self.world > 1:
                dist.all_reduce(box, op=dist.ReduceOp.MAX)
            box = box.cpu().numpy()
            lo, hi = -box[self.axis], box[3 + self.axis]
            hi = hi + max(1e-6 * abs(hi - lo), 1e-30)
            # 2. balanced cuts
            self._histogram(points, lo, hi)
            if self.world > 1:
                dist.all_reduce(self._hist, op=dist.ReduceOp.SUM)
            hist = self._hist.cpu().numpy()
            self.n_global = int(hist.sum())
            self.cuts = balanced_cuts(hist, lo, hi, self.world, cost_weights(hist) if self.balance == "cost" else None)
            # what this rank should own under these cuts: the yardstick of the re-balance request below
            edges = np.clip(np.round((self.cuts[1:-1].astype(np.float64) - float(lo)) / ((float(hi) - float(lo)) / self.n_bins)).astype(np.int64), 0, self.n_bins)
            edges = np.concatenate([[0], edges, [self.n_bins]])
            self._expected_owned = int(hist[edges[self.rank]:edges[self.rank + 1]].sum())
            self.n_recuts += 1
        # 3. partition, 4. exchange (cut coordinates are open ended at both ends, so points that left the old box still have an owner)
        want = 0
        if self.n_global > 0 and self.n_owned > 0:
            mean = self.n_global / self.world
            expect = self._expected_owned if self._expected_owned is not None else mean
            want = int(abs(self.n_owned - expect) > self.rebalance_tolerance * mean)     # judged on the previous step's balance
        if self.exchange == "p2p":
            self.local, self.n_owned, self.n_halo, flag = self._push_exchange(points, id_base, want)
        else:
            counts = self._partition(points, id_base, self.cuts, halo_width(self.radius))
            self.local, self.n_owned, self.n_halo, flag = exchange_records(dist, self._records, counts, self.world, want)
        self._recut = bool(flag)
        # 5. local search, halo points find-only
        eng = self.engine
        eng.set_option(L.TNSB_OPT_QUERY_LIMIT, self.n_owned)
        if not self._set_added:
            eng.add_point_set(self.local, n_points=self.local.shape[0])
            eng.set_active_search(0, 0, True)
            self._set_added = True
        else:
            eng.resize_point_set(0, self.local, n_points=self.local.shape[0])
        eng.run()

    # ---- results
    def local_ids(self) -> np.ndarray:
        """Global id of every local point ([owned | halo] order)."""
        return self.local[:, 3].contiguous().view(self.torch.int32).cpu().numpy()

    def owned_lists_global(self):
        """{global id of owned point: sorted array of global neighbour ids} -- test helper, small inputs only."""
        ids = self.local_ids()
        ragged, pos = self.engine.neighbor_lists(0, 0)
        out = {}
        for i in range(self.n_owned):
            p = int(pos[i])
            n = int(ragged[p])
            out[int(ids[i])] = np.sort(ids[ragged[p + 1: p + 1 + n]])
        return out


class ShardedUniformJob:
    """bench.py helper: this rank's 1/world chunk of a synthetic cloud, stepped device-resident or end-to-end."""

    def __init__(self, workload, points_per_gpu, rank, world, local_rank, stream, shard_input="slab"):
        import torch
        import torch.distributed as dist
        from . import clouds
        total = points_per_gpu * world
        if workload not in ("uniform", "clustered"):
            raise SystemExit("multi-GPU bench supports the uniform and the clustered workload")
        self.radius = float(clouds.radius_for_mean_neighbors(total))
        chunk = clouds.uniform_cloud(points_per_gpu, 42 + rank)          # i.i.d. uniform chunk of the global cloud
        if workload == "clustered":
            # density gradient along the slab axis: z -> z^1.5 of the GLOBAL coordinate (the density grows like z^(-1/3) towards z = 0, the
            # count-balanced slabs get unequal thickness, the lowest slab holds the dense end with its long lists).  z is kept >= 1e-3
            # (44x the mean density at most) so that the workload stays a clustered cloud and not a stress test of a singularity.
            chunk = chunk.copy()
            zg = (chunk[:, 2].astype(np.float64) + rank) / world if shard_input == "slab" else chunk[:, 2].astype(np.float64)
            chunk[:, 2] = (1e-3 + (1.0 - 1e-3) * zg ** 1.5).astype(np.float32)
        elif shard_input == "slab":
            # the cloud is ALREADY sharded by Z slab, as in a running simulation: rank r holds z in [r/world, (r+1)/world).  A step then
            # exchanges the one-cell halo plus the few points that the count-balanced cuts move across a boundary.  "random" hands
            # every rank an i.i.d. sample of the whole cube instead: (world-1)/world of all points change rank in every step.
            chunk = chunk.copy()
            chunk[:, 2] = (chunk[:, 2] + np.float32(rank)) / np.float32(world)
        self.h_pts = torch.from_numpy(chunk).pin_memory()
        self.d_pts = self.h_pts.cuda(non_blocking=False)
        self.d_stage = torch.empty_like(self.d_pts)
        self.id_base = rank * points_per_gpu
        self.search = ShardedSearch(self.radius, rank, world, local_rank, stream=stream, dist=dist if world > 1 else None)
        self._stats = None
        self._stats_e2e = None

    def step_device(self):
        self.search.engine.set_option(L.TNSB_OPT_HOST_RESULTS, 0)
        self.search.step(self.d_pts, self.id_base)
        self._stats = self.search.engine.stats()

    def step_e2e(self):
        self.search.engine.set_option(L.TNSB_OPT_HOST_RESULTS, 1)
        self.d_stage.copy_(self.h_pts, non_blocking=True)               # host -> device of this step's inputs
        self.search.step(self.d_stage, self.id_base)
        st = self.search.engine.stats()
        st["h2d_bytes"] = int(self.h_pts.numel() * 4)
        self._stats_e2e = st

    def stats(self):
        return self._stats

    def stats_e2e(self):
        return self._stats_e2e
