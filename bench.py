#!/usr/bin/env python
"""bench.py -- million neighbour-queries/s (build + query) of the B200 engine, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points-per-gpu P]
                    [--workload uniform|dambreak|twoset] [--quick] [--no-cpu-baseline] [--shard-input slab|random]

A step is one full pass of the hot path (everything tns::TreeNSearch::run() does: world box, cell keys, bucket sort, reorder,
cell table, distance query, neighbour lists) over one synthetic cloud.

  N = 1 : BASELINE.json configs[1]: 10M uniform-random points, single set, fixed radius (k_mean ~ 29.7).  The same line also
          carries (unless --quick): config 3 (10M dam-break, zsort every 10 of 20 steps), config 4 (2M + 500K, variable radii),
          the 80M cloud on ONE GPU (the denominator of the strong-scaling claim) and the reference's own micro-benchmark
          (9 261 lattice points, tests/tests.cpp:239-279), each with the reference CPU timing beside it.
  N > 1 : the same density with 10M points PER GPU (weak scaling; N = 8 is configs[4], 80M points), Z-slab sharded with a
          one-cell halo exchanged over NVLink (treensearch_b200/sharded.py); rank 0 also times the WHOLE cloud on one GPU
          (`strong_scaling`), a second arm redistributes i.i.d. chunks (`random_input`), and a parity gate compares neighbour
          totals and list digests of the sharded run with the single-GPU run.

`value`  : inputs resident in HBM, lists left in HBM, timed with CUDA events on the stream the kernels run on.
`e2e`    : the same metric through the drop-in C++ class (include/TreeNSearch) with a PAGEABLE std::vector of points in and
           host-addressable lists out (tools/bench_cpp.cpp); the Python mirror with pinned tensors is reported beside it.
`--impl reference` : the unmodified reference's run() (oracle/_ref, AVX2 + OpenMP, ALL host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "million neighbor-queries/sec (build+query)"
UNIT = "Mq/s"
MAX_REFERENCE_POINTS = 80_000_000        # the reference arm runs the stated workload up to this size (80M: ~12 GB of lists)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def make_cloud(workload, n_points, seed=42):
    from treensearch_b200 import clouds
    if workload == "uniform":
        pts = clouds.uniform_cloud(n_points, seed)
        r = float(clouds.radius_for_mean_neighbors(n_points))
    elif workload == "dambreak":
        pts, _d, r = clouds.dam_break_cloud(n_points)
        r = float(r)
    else:
        raise SystemExit(f"unknown workload {workload}")
    return pts, r


def sharded_global_cloud(points_per_gpu, world, shard_input="slab"):
    """The global cloud of the N-GPU arm, in global id order (rank r contributes ids [r*P, (r+1)*P)): what treensearch_b200.sharded
    .ShardedUniformJob builds rank by rank."""
    from treensearch_b200 import clouds
    chunks = []
    for rank in range(world):
        chunk = clouds.uniform_cloud(points_per_gpu, 42 + rank)
        if shard_input == "slab":
            chunk = chunk.copy()
            chunk[:, 2] = (chunk[:, 2] + np.float32(rank)) / np.float32(world)
        chunks.append(chunk)
    pts = np.concatenate(chunks) if world > 1 else chunks[0]
    return np.ascontiguousarray(pts), float(clouds.radius_for_mean_neighbors(points_per_gpu * world))


WORKLOAD_NAMES = {
    "uniform": "uniform-random points in the unit cube, single set, fixed radius (k_mean~29.7)",
    "dambreak": "SPH dam-break clustered points (~60 neighbours interior), single set, fixed radius, zsort every 10 of 20 steps",
    "twoset": "two point sets (2M fluid + 500K boundary), searches 0->0 0->1 1->0, variable per-point radii, symmetric search",
}


def workload_config(args, total):
    return {"workload": f"{total} {WORKLOAD_NAMES[args.workload]}" + (f", Z-slab sharded over {args.gpus} GPUs with a one-cell halo exchange over NVLink" if args.gpus > 1 else ", 1xB200"),
            "n_points": total, "points_per_gpu": args.points_per_gpu, "active_searches": "0->0" if args.workload != "twoset" else "0->0, 0->1, 1->0",
            **({"shard_input": {"slab": "every rank holds its own Z slab (halo + migrating points exchanged per step)",
                                "random": "every rank holds an i.i.d. sample of the whole cube (full redistribution per step)"}[args.shard_input]}
               if args.gpus > 1 else {}),
            "l2": "flushed between timed steps (256 MiB write, untimed); per-step working set ~1.9 GB >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------ reference arm
def _reference_threads():
    """All host cores for the reference's OpenMP runtime -- torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, so the
    variable is overridden BEFORE libgomp is loaded and the thread count is set explicitly on every handle as well."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return cores


def time_reference(pts, r, steps, warmup, unsorted_too=True):
    """The unmodified reference on the host cores: zsort once (its intended regime, README.md:142-143), then time run()."""
    cores = _reference_threads()
    from oracle import loader
    ref = loader.Reference()
    ref.set_n_threads(cores)
    pts = np.ascontiguousarray(pts).copy()
    ref.set_search_radius(r)
    ref.add_point_set(pts)
    ref.set_active_search(0, 0, True)
    ms_unsorted = None
    if unsorted_too:
        ref.time_runs(1)                        # raw (unsorted) input: one warm-up + one timed run
        ms_unsorted = float(ref.time_runs(1)[0])
    t0 = time.time()
    ref.prepare_zsort()
    ref.apply_zsort(0, pts.reshape(-1), 3)
    ms_zsort = (time.time() - t0) * 1e3
    ref.time_runs(max(warmup, 1))
    ms = ref.time_runs(steps)
    n = pts.shape[0]
    out = {
        "ms_per_step": float(np.mean(ms)), "ms_best": float(np.min(ms)), "ms_zsort": ms_zsort,
        "value": n / (float(np.mean(ms)) * 1e-3) / 1e6, "cores": cores, "omp_max_threads": loader.Reference.n_threads(), "n_points": n,
    }
    if ms_unsorted is not None:
        out["ms_unsorted_input"] = ms_unsorted
        out["value_unsorted_input"] = n / (ms_unsorted * 1e-3) / 1e6
    ref.close()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _reference_threads()
    from oracle import loader
    if not loader.reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtns_ref.so missing (make -C oracle where /root/reference exists)"}))
        return
    total = args.points_per_gpu * args.gpus
    if args.workload == "uniform":
        n = min(total, MAX_REFERENCE_POINTS)
        if args.gpus > 1 and n == total:
            pts, r = sharded_global_cloud(args.points_per_gpu, args.gpus, args.shard_input)      # the very cloud the N-GPU arm searches
        else:
            pts, r = make_cloud("uniform", n)
    elif args.workload == "dambreak":
        n = min(total, MAX_REFERENCE_POINTS)
        pts, r = make_cloud("dambreak", n)
    else:
        raise SystemExit("--impl reference supports the uniform and dambreak workloads")
    t = time_reference(pts, r, args.steps, args.warmup, unsorted_too=(n <= 10_000_000))
    sample = (f"the stated workload: all {n} points" if n == total else f"SAMPLE: {n} of {total} points, same k_mean") + \
             f" ({args.workload}), z-sorted input, {args.steps} run() calls after {max(args.warmup, 1)} warm-up, {t['cores']} OpenMP threads"
    cfg = workload_config(args, total)
    if n != total:
        cfg["workload"] = f"SAMPLE of {n} points of: " + cfg["workload"]
    line = {
        "metric": METRIC, "value": t["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference", "config": cfg,
        "cpu_baseline": {"value": t["value"], "unit": UNIT, "cores": t["cores"], "kind": "reference", "sample": sample,
                         "omp_max_threads": t["omp_max_threads"], "ms_zsort_once": t["ms_zsort"],
                         **({"value_unsorted_input": t["value_unsorted_input"]} if "value_unsorted_input" in t else {})},
        "e2e": {"value": t["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------ helpers of our arm
class Timer:
    """K timed steps with CUDA events on `stream`, L2 flushed before every step, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, torch, stream, dist):
        self.torch, self.stream, self.dist = torch, stream, dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, step, k, w, between=None):
        torch = self.torch
        for _ in range(w):
            step()
            if between:
                between()
        self.barrier()
        ms = []
        for _ in range(k):
            self.flush.fill_(1)                      # L2 flush, outside the timed pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if self.dist is not None:
                self.dist.barrier()                  # ranks start every timed step together (the flush / sync above de-synchronises them)
            torch.cuda.synchronize()
            e0.record(self.stream)
            step()
            e1.record(self.stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
            if between:
                between()
        self.barrier()
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(tot, op=self.dist.ReduceOp.MAX)      # max over ranks
        return float(tot.item()) / k, ms


def pcie_rates(torch):
    """Pinned host <-> device copy rates of this box (GB/s): the floor of any host-in / host-out arm."""
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    out = {}
    for name, (dst, src) in {"h2d_gbs": (d, h), "d2h_gbs": (h, d)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        e1.record()
        e1.synchronize()
        out[name] = 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return out


def single_gpu_engine(t, stream, d_pts, r, host_results=False):
    eng = t.TreeNSearch(d_pts.device.index if hasattr(d_pts, "device") and d_pts.is_cuda else -1)
    eng.set_stream(stream.cuda_stream)
    if not host_results:
        eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
    eng.set_search_radius(r)
    eng.add_point_set(d_pts)
    eng.set_active_search(0, 0, True)
    return eng


def device_list_digests(torch, ragged, list_pos, sample_idx, id_map=None):
    """Order independent digest (wrapping sum of a 64-bit mix of the neighbour ids, + count) of the lists of `sample_idx`, computed on
    the device from the engine's ragged buffer.  id_map: local -> global id (None: identity)."""
    pos = list_pos[sample_idx]
    cnt = ragged[pos].to(torch.int64)
    starts = pos + 1
    total = int(cnt.sum().item())
    owner = torch.repeat_interleave(torch.arange(sample_idx.shape[0], device=ragged.device), cnt)
    first = torch.cumsum(cnt, 0) - cnt
    within = torch.arange(total, device=ragged.device) - first[owner]
    ids = ragged[starts[owner] + within].to(torch.int64)
    if id_map is not None:
        ids = id_map[ids].to(torch.int64)
    h = ids * -7046029254386353131                   # 0x9E3779B97F4A7C15 as int64, wrapping multiply
    h = h ^ (h >> 29)
    dig = torch.zeros(sample_idx.shape[0], dtype=torch.int64, device=ragged.device)
    dig.index_add_(0, owner, h)
    return dig + cnt


def engine_device_lists(torch, eng, n_lists):
    import ctypes as C
    d_ragged, d_pos, n_ints = eng.neighbor_lists_device(0, 0)

    class _Arr:
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}
    ragged = torch.as_tensor(_Arr(d_ragged, max(n_ints, 1), "<i4"), device="cuda")
    pos = torch.as_tensor(_Arr(d_pos, max(n_lists, 1), "<i8"), device="cuda")
    return ragged, pos


# ------------------------------------------------------------------------------------------------------------ extra workloads (N = 1)
def bench_c3_dambreak(t, torch, timer, stream, n, with_cpu):
    """Config 3: 10M dam-break points, prepare_zsort + apply_zsort before steps 0 and 10 of 20, points advected (on the device, untimed)
    between steps; ms per step = mean run() + the zsort cost amortised over the loop."""
    from treensearch_b200 import clouds
    pts_np, d, r = clouds.dam_break_cloud(n)
    d_pts = torch.from_numpy(pts_np).cuda()
    eng = single_gpu_engine(t, stream, d_pts, float(r))
    a = 0.1 * float(d) / np.sqrt(2.0)
    state = {"step": 0}

    def advect():
        ph = 0.37 * state["step"]
        p = d_pts.to(torch.float64)
        u = torch.stack([a * torch.sin(2.0 * p[:, 1] + ph) * torch.cos(3.0 * p[:, 2]),
                         a * torch.sin(2.0 * p[:, 2] + ph) * torch.cos(3.0 * p[:, 0]),
                         a * torch.sin(2.0 * p[:, 0] + ph) * torch.cos(3.0 * p[:, 1])], dim=1)
        d_pts.copy_((p + u).to(torch.float32))
        state["step"] += 1

    def zsort():
        eng.prepare_zsort()
        eng.apply_zsort(0, d_pts, 3)

    eng.run(); zsort(); eng.run()                     # warm-up (buffers of run() and of the zsort path, column height)
    ms_run, ms_zs = [], []
    for step in range(20):
        if step % 10 == 0:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            zsort()
            torch.cuda.synchronize()
            ms_zs.append((time.perf_counter() - t0) * 1e3)
        ms, _ = timer.run(eng.run, 1, 0)
        ms_run.append(ms)
        advect()
    st = eng.stats()
    ms_step = float(np.mean(ms_run)) + sum(ms_zs) / 20.0
    out = {"config": f"{n} {WORKLOAD_NAMES['dambreak']}", "ms_run_mean": float(np.mean(ms_run)), "ms_zsort_prepare_plus_apply": float(np.mean(ms_zs)),
           "ms_per_step": ms_step, "value": n / (ms_step * 1e-3) / 1e6, "unit": UNIT, "neighbors_per_query": st["n_neighbors"] / max(st["n_queries"], 1),
           "stages_ms": {k: st[k] for k in ("ms_aabb", "ms_keys", "ms_sort", "ms_reorder", "ms_cells", "ms_query", "ms_total_device")},
           "brick_query": st["brick_query"], "n_slow_queries": st["n_slow_queries"]}
    if with_cpu:
        tr = time_reference(pts_np, float(r), 3, 1, unsorted_too=False)
        ref_step = tr["ms_per_step"] + 2.0 * tr["ms_zsort"] / 20.0
        out["cpu_baseline"] = {"value": n / (ref_step * 1e-3) / 1e6, "unit": UNIT, "cores": tr["cores"], "kind": "reference", "ms_run_mean": tr["ms_per_step"],
                               "ms_zsort_prepare_plus_apply": tr["ms_zsort"], "ms_per_step": ref_step,
                               "sample": f"all {n} points, zsort cost amortised the same way (2 of 20 steps), mean of 3 run() after 1 warm-up"}
    eng.close()
    return out


def bench_c4_twoset(t, torch, timer, stream, with_cpu):
    """Config 4: 2M fluid + 500K boundary points, per-point radii, searches 0->0, 0->1, 1->0, symmetric search on (default) and off."""
    from treensearch_b200 import clouds
    p0, r0, p1, r1, _ = clouds.two_set_cloud(2_000_000, 500_000)
    d = [torch.from_numpy(x).cuda() for x in (p0, r0, p1, r1)]
    n_search = p0.shape[0] + p1.shape[0]
    out = {"config": f"{WORKLOAD_NAMES['twoset']}", "searching_points": n_search}
    for sym in (True, False):
        eng = t.TreeNSearch(0)
        eng.set_stream(stream.cuda_stream)
        eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
        eng.add_point_set(d[0], d[1])
        eng.add_point_set(d[2], d[3])
        for pr in ((0, 0), (0, 1), (1, 0)):
            eng.set_active_search(*pr, True)
        eng.set_symmetric_search(sym)
        ms, _ = timer.run(eng.run, 5, 3)
        st = eng.stats()
        key = "symmetric" if sym else "asymmetric"
        out[key] = {"ms_per_step": ms, "value": n_search / (ms * 1e-3) / 1e6, "unit": UNIT, "n_pair_queries": st["n_queries"],
                    "neighbors_per_pair_query": st["n_neighbors"] / max(st["n_queries"], 1), "ms_query": st["ms_query"], "n_slow_queries": st["n_slow_queries"]}
        eng.close()
        if with_cpu:
            cores = _reference_threads()
            from oracle import loader
            ref = loader.Reference()
            ref.set_n_threads(cores)
            q0, q1 = p0.copy(), p1.copy()
            s0, s1 = r0.copy(), r1.copy()
            ref.add_point_set(q0, s0)
            ref.add_point_set(q1, s1)
            for pr in ((0, 0), (0, 1), (1, 0)):
                ref.set_active_search(*pr, True)
            ref.set_symmetric_search(sym)
            ref.prepare_zsort()
            for s, (pp, rr) in enumerate(((q0, s0), (q1, s1))):
                ref.apply_zsort(s, pp.reshape(-1), 3)
                ref.apply_zsort(s, rr, 1)
            ref.time_runs(1)
            mr = float(np.mean(ref.time_runs(3)))
            out[key]["cpu_baseline"] = {"value": n_search / (mr * 1e-3) / 1e6, "unit": UNIT, "cores": cores, "kind": "reference", "ms_per_step": mr,
                                        "sample": "the stated workload, z-sorted input, mean of 3 run() after 1 warm-up"}
            ref.close()
    return out


def bench_small_n(t, torch, stream, with_cpu):
    """The reference's own micro-benchmark (tests/tests.cpp:239-279): 9 261 lattice points, z-sorted, 1000 x run(), wall clock per call."""
    from treensearch_b200 import clouds
    pts, r = clouds.sph_lattice(9000)              # benchmark_one_dynamic_set(9000): a 21^3 = 9261 point lattice
    pts = pts.copy()
    out = {"config": f"{pts.shape[0]} lattice points (tests/tests.cpp:239-279), z-sorted, 1000 run() calls, host wall clock per call"}
    # the reference's OWN test program, unmodified, once linked to this engine through the drop-in header and once built with the
    # reference's own library: the "Runtime parallel SIMD" line both print after 1000 x run() (oracle/Makefile: ref-tests)
    from oracle import loader as _ld
    import re
    import subprocess
    for key, exe in (("cpp_program_on_b200_ms_per_run", _ld.REF_TESTS_BIN), ("cpp_program_reference_ms_per_run", getattr(_ld, "REF_TESTS_NATIVE_BIN", ""))):
        if not (with_cpu and exe and os.path.exists(exe)):
            continue
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(_reference_threads())
        try:
            res = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
            m = re.search(r"Runtime parallel SIMD: ([0-9.eE+-]+) ms", res.stdout)
            if m:
                out[key] = float(m.group(1))
        except Exception as e:      # noqa: BLE001 -- a missing / failing helper binary must not take the bench line down
            out[key + "_error"] = str(e)[:200]
    for name, host in (("host_arrays_ms_per_run", True), ("device_resident_ms_per_run", False)):
        arr = pts.copy() if host else torch.from_numpy(pts).cuda()
        eng = t.TreeNSearch(0)
        eng.set_stream(stream.cuda_stream)
        if not host:
            eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
        eng.set_search_radius(float(r))
        eng.add_point_set(arr)
        eng.set_active_search(0, 0, True)
        eng.prepare_zsort()
        eng.apply_zsort(0, arr, 3)
        for _ in range(20):
            eng.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(1000):
            eng.run()
        torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t0)
        out[name.replace("_ms_per_run", "_graph_replay")] = int(eng.stats()["graph_replay"])
        eng.close()
    if with_cpu:
        cores = _reference_threads()
        from oracle import loader
        ref = loader.Reference()
        ref.set_n_threads(cores)
        p = pts.copy()
        ref.set_search_radius(float(r))
        ref.add_point_set(p)
        ref.set_active_search(0, 0, True)
        ref.prepare_zsort()
        ref.apply_zsort(0, p.reshape(-1), 3)
        ref.time_runs(5)
        out["reference_ms_per_run"] = float(np.mean(ref.time_runs(1000)))
        out["reference_cores"] = cores
        ref.close()
    return out


def bench_cpp_e2e(n, steps, warmup):
    """tools/bench_cpp.cpp: the drop-in C++ class with a pageable std::vector, default options and with TNSB_OPT_PIN_USER_MEMORY."""
    from treensearch_b200 import build
    exe = build.BENCH_CPP
    if not os.path.exists(exe):
        exe = build.build_bench_cpp()
    out = {}
    for pin in (0, 1):
        res = subprocess.run([exe, str(n), str(steps), str(warmup), str(pin)], capture_output=True, text=True, timeout=600)
        if res.returncode != 0:
            out["pinned_registered" if pin else "pageable"] = {"error": (res.stdout + res.stderr)[-300:]}
            continue
        out["pinned_registered" if pin else "pageable"] = json.loads(res.stdout.strip().splitlines()[-1])
    return out


# ------------------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import treensearch_b200 as t

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; treensearch_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=30))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if world > 1 and args.workload != "uniform":
        raise SystemExit("the multi-GPU bench supports the uniform workload")

    total = args.points_per_gpu * args.gpus
    stream = torch.cuda.current_stream()
    timer = Timer(torch, stream, dist)
    with_cpu = not args.no_cpu_baseline
    extras = {}

    if world == 1:
        if args.workload == "twoset":
            res = bench_c4_twoset(t, torch, timer, stream, with_cpu)
            print(json.dumps({"metric": METRIC, "unit": UNIT, "n_gpus": 1, "value": res["symmetric"]["value"], "ms_per_step": res["symmetric"]["ms_per_step"],
                              "higher_is_better": True, "dtype": "f32", "data": "synthetic", "config": {"workload": res["config"]}, "twoset": res}))
            return
        pts_np, r = make_cloud(args.workload, total)
        # ---- device-resident arm
        d_pts = torch.from_numpy(pts_np).cuda()
        eng = single_gpu_engine(t, stream, d_pts, r)
        step_dev = eng.run
        # ---- Python mirror end to end: pinned host points in, host-addressable lists out
        h_pts = torch.from_numpy(pts_np).pin_memory()
        eng_e2e = single_gpu_engine(t, stream, h_pts, r, host_results=True)
        step_e2e = eng_e2e.run
        stats_of, stats_e2e = eng.stats, eng_e2e.stats
    else:
        from treensearch_b200 import sharded
        job = sharded.ShardedUniformJob(args.workload, args.points_per_gpu, rank, world, local_rank, stream, args.shard_input)
        r = job.radius
        step_dev, step_e2e = job.step_device, job.step_e2e
        stats_of, stats_e2e = job.stats, job.stats_e2e

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, ms_list = timer.run(step_dev, args.steps, max(args.warmup, 3))
    st = stats_of()
    ms_e2e, _ = timer.run(step_e2e, max(1, min(args.steps, 5)), 3)
    st2 = stats_e2e()
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (device-resident and end-to-end)

    # ---- N > 1: parity gate, redistribution arm, strong-scaling denominator
    if world > 1:
        from treensearch_b200 import sharded
        # (1) redistribution arm: every rank starts from an i.i.d. chunk of the whole cube
        other = "random" if args.shard_input == "slab" else "slab"
        job2 = sharded.ShardedUniformJob(args.workload, args.points_per_gpu, rank, world, local_rank, stream, other)
        ms_other, _ = timer.run(job2.step_device, max(1, min(args.steps, 5)), 3)
        extras[f"{other}_input"] = {"ms_per_step": ms_other, "value": total / (ms_other * 1e-3) / 1e6, "unit": UNIT,
                                    "note": "every rank holds an i.i.d. sample of the cube: a step redistributes (N-1)/N of all points" if other == "random" else "every rank holds its own Z slab"}
        del job2
        # (1b) clustered arm (north_star: "uniform and clustered clouds at 1/2/4/8 GPUs"): density gradient z -> z^1.5 along the slab axis
        if args.workload == "uniform":
            try:
                job3 = sharded.ShardedUniformJob("clustered", args.points_per_gpu, rank, world, local_rank, stream, "slab")
                ms_cl, _ = timer.run(job3.step_device, max(1, min(args.steps, 5)), 3)
                st3 = job3.stats()
                cnt3 = torch.tensor([st3["n_neighbors"], job3.search.n_owned, st3["n_slow_queries"]], dtype=torch.int64, device="cuda")
                dist.all_reduce(cnt3, op=dist.ReduceOp.SUM)
                extras["clustered_z15"] = {"ms_per_step": ms_cl, "value": total / (ms_cl * 1e-3) / 1e6, "unit": UNIT,
                                           "neighbors_per_query": float(cnt3[0].item()) / max(float(cnt3[1].item()), 1.0), "n_slow_queries": int(cnt3[2].item()),
                                           "note": "same point count, z -> 1e-3 + z^1.5: density gradient along the slab axis (up to 44x the mean at the dense end), "
                                                   "slab cuts weighted by the estimated cost per histogram bin (sharded.cost_weights: the dense end owns fewer points), same radius as the uniform arm"}
                del job3
            except Exception as e:      # noqa: BLE001
                extras["clustered_z15"] = {"error": repr(e)[:300]}
        # (2) parity gate + strong scaling: rank 0 searches the WHOLE cloud on one GPU; neighbour totals and list digests of a sample of
        #     every rank's owned points must agree with it
        job.step_device()
        torch.cuda.synchronize()
        S = job.search
        n_owned = S.n_owned
        ragged, pos = engine_device_lists(torch, S.engine, n_owned)
        local_ids = S.local[:, 3].contiguous().view(torch.int32)
        k = min(n_owned, 50_000)
        sample = torch.arange(k, device="cuda") * max(n_owned // max(k, 1), 1)
        sample = sample[sample < n_owned]
        dig = device_list_digests(torch, ragged, pos, sample, id_map=local_ids)
        gids = local_ids[sample].to(torch.int64)
        cnt_local = torch.tensor([S.engine.stats()["n_neighbors"], n_owned], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt_local, op=dist.ReduceOp.SUM)
        pad = torch.full((50_000, 2), -1, dtype=torch.int64, device="cuda")
        pad[: gids.shape[0], 0], pad[: gids.shape[0], 1] = gids, dig
        gathered = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, gathered, dst=0)
        if rank == 0:
            pts_all, r_all = sharded_global_cloud(args.points_per_gpu, world, args.shard_input)
            d_all = torch.from_numpy(pts_all).cuda()
            del pts_all
            eng1 = single_gpu_engine(t, stream, d_all, r_all)
            t1 = Timer(torch, stream, None)
            ms_1, _ = t1.run(eng1.run, 3, 2)
            st1 = eng1.stats()
            rag1, pos1 = engine_device_lists(torch, eng1, total)
            allg = torch.cat(gathered)
            allg = allg[allg[:, 0] >= 0]
            dig1 = device_list_digests(torch, rag1, pos1, allg[:, 0])
            bad = int((dig1 != allg[:, 1]).sum().item())
            extras["parity_gate"] = {"neighbors_sharded": int(cnt_local[0].item()), "neighbors_single_gpu": int(st1["n_neighbors"]),
                                     "owned_points": int(cnt_local[1].item()), "sampled_lists": int(allg.shape[0]), "sampled_lists_differing": bad,
                                     "ok": bool(bad == 0 and int(cnt_local[0].item()) == int(st1["n_neighbors"]) and int(cnt_local[1].item()) == total)}
            extras["strong_scaling"] = {"single_gpu_same_cloud": {"n_points": total, "ms_per_step": ms_1, "value": total / (ms_1 * 1e-3) / 1e6, "unit": UNIT,
                                                                   "stages_ms": {k2: st1[k2] for k2 in ("ms_reorder", "ms_query", "ms_total_device")}},
                                        "n_gpus": world, "ms_per_step": ms_dev, "strong_speedup": ms_1 / ms_dev,
                                        "note": "the SAME cloud on one GPU vs sharded over N GPUs (north_star: >= 6x at 8 GPUs on 80M points)"}
            eng1.close()
            del d_all
        dist.barrier()

    # ---- roofline of the dominant kernel (the distance query): algorithmic bytes per launch / CUDA-event duration
    # query = 24 B per point + 4 B per neighbour id (read sorted xyz 12 + idx 4, write count 4 + offset 4, write k ids)
    peak, peak_src = measured_peak_gbs()
    q_bytes = 24.0 * st["n_queries"] + 4.0 * st["n_neighbors"]
    q_ms = st["ms_query"] / max(st["n_query_launches"], 1)
    achieved = q_bytes / max(st["n_query_launches"], 1) / (q_ms * 1e-3) / 1e9 if q_ms > 0 else 0.0
    run_bytes = 92.0 * st["n_queries"] + 4.0 * st["n_neighbors"]          # whole run(): B_alg = 92 + 4k bytes per query
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "query_kernel_dram_bytes.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = total / (ms_dev * 1e-3) / 1e6
    e2e_py = total / (ms_e2e * 1e-3) / 1e6
    kernel_name = "brick_query_kernel (half-radius grid, lane = query)" if st.get("brick_query") else "query_kernel (27-cell query)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, total),
        "gpu_launches": int(st["n_kernel_launches"]) * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms": q_ms, "algorithmic_bytes_per_launch": q_bytes / max(st["n_query_launches"], 1),
                     "whole_run": {"algorithmic_bytes": run_bytes, "ms": st["ms_total_device"],
                                   "achieved": run_bytes / (st["ms_total_device"] * 1e-3) / 1e9 if st["ms_total_device"] > 0 else 0.0,
                                   "frac": run_bytes / (st["ms_total_device"] * 1e-3) / 1e9 / peak if st["ms_total_device"] > 0 else 0.0}},
        "stages_ms": {k: st[k] for k in ("ms_aabb", "ms_keys", "ms_sort", "ms_reorder", "ms_cells", "ms_query", "ms_total_device")},
        "neighbors_per_query": st["n_neighbors"] / max(st["n_queries"], 1),
        "n_slow_queries": st.get("n_slow_queries", 0),
        "step_ms_all": ms_list,
    }
    e2e_python = {"value": e2e_py, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(st2["h2d_bytes"]) * world,
                  "d2h_bytes_per_step": int(st2["d2h_bytes"]) * world,
                  "stages_ms": {k: st2[k] for k in ("ms_upload", "ms_total_device", "ms_download", "ms_wall")},
                  "note": "Python mirror, PINNED host points in; count-prefixed lists written by the query kernel into mapped pinned host memory (zero-copy) + list_pos copied back"
                          + ("; bytes = rank 0 x n_gpus" if world > 1 else "")}
    rates = pcie_rates(torch)
    floor_ms = (e2e_python["h2d_bytes_per_step"] / world / (rates["h2d_gbs"] * 1e9) + e2e_python["d2h_bytes_per_step"] / world / (rates["d2h_gbs"] * 1e9)) * 1e3
    if world == 1 and args.workload == "uniform" and not args.quick:
        try:
            cpp = bench_cpp_e2e(total, max(1, min(args.steps, 5)), 2)
        except Exception as e:                        # the line must still be printed
            cpp = {"error": repr(e)[:200]}
    else:
        cpp = None
    if cpp and "pageable" in cpp and "ms_mean" in cpp["pageable"]:
        pg = cpp["pageable"]
        line["e2e"] = {"value": total / (pg["ms_mean"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": pg["ms_mean"],
                       "h2d_bytes_per_step": pg["h2d_bytes"], "d2h_bytes_per_step": pg["d2h_bytes"],
                       "floor_ms": floor_ms, "pcie_measured": rates,
                       "note": "drop-in C++ class (include/TreeNSearch, tools/bench_cpp.cpp): PAGEABLE std::vector<float> in, run(), host-addressable "
                               "neighbour lists out, host wall clock; floor_ms = this step's PCIe bytes / the measured pinned copy rates",
                       "cpp": cpp, "python_pinned": e2e_python}
    else:
        line["e2e"] = dict(e2e_python, floor_ms=floor_ms, pcie_measured=rates, **({"cpp": cpp} if cpp else {}))
    line.update(extras)

    # ---- CPU baseline beside it (rank 0, N = 1 only): the unmodified reference on the same workload
    if world == 1 and with_cpu:
        from oracle import loader
        if loader.reference_available():
            tr = time_reference(pts_np, r, 3, 1)
            line["cpu_baseline"] = {"value": tr["value"], "unit": UNIT, "cores": tr["cores"], "kind": "reference",
                                    "sample": f"all {total} points of the same workload, z-sorted input, mean of 3 run() after 1 warm-up",
                                    "ms_per_step": tr["ms_per_step"], "value_unsorted_input": tr.get("value_unsorted_input"), "omp_max_threads": tr["omp_max_threads"]}
        else:
            from treensearch_b200 import clouds
            n = 1_000_000
            pts_cpu = clouds.uniform_cloud(n, 42)
            port = loader.OraclePort()
            port.set_search_radius(float(clouds.radius_for_mean_neighbors(n)))
            port.add_point_set(pts_cpu)
            port.set_active_search(0, 0, True)
            t0 = time.time()
            port.run(1)
            dt = time.time() - t0
            line["cpu_baseline"] = {"value": n / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n} uniform points, restated grid oracle (oracle/_ref not present)"}

    # ---- the other configurations of BASELINE.json and the strong-scaling denominator, as extra keys of the N = 1 line
    if world == 1 and args.workload == "uniform" and not args.quick:
        eng.close(); eng_e2e.close()
        del d_pts, h_pts
        torch.cuda.empty_cache()
        ref_ok = with_cpu
        try:
            from oracle import loader
            ref_ok = with_cpu and loader.reference_available()
        except Exception:
            ref_ok = False
        for key, fn in (("dambreak_c3", lambda: bench_c3_dambreak(t, torch, timer, stream, 10_000_000, ref_ok)),
                        ("twoset_c4", lambda: bench_c4_twoset(t, torch, timer, stream, ref_ok)),
                        ("small_n_latency", lambda: bench_small_n(t, torch, stream, ref_ok))):
            try:
                line[key] = fn()
            except Exception as e:
                line[key] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        try:
            n80 = 80_000_000
            pts80, r80 = sharded_global_cloud(10_000_000, 8, "slab")
            d80 = torch.from_numpy(pts80).cuda()
            del pts80
            e80 = single_gpu_engine(t, stream, d80, r80)
            ms80, _ = timer.run(e80.run, 3, 2)
            s80 = e80.stats()
            line["strong_80m_1gpu"] = {"n_points": n80, "ms_per_step": ms80, "value": n80 / (ms80 * 1e-3) / 1e6, "unit": UNIT,
                                       "stages_ms": {k: s80[k] for k in ("ms_aabb", "ms_keys", "ms_sort", "ms_reorder", "ms_cells", "ms_query", "ms_total_device")},
                                       "note": "BASELINE configs[4]'s 80M cloud on ONE GPU: the denominator of the >= 6x strong-scaling target at 8 GPUs"}
            e80.close()
            del d80
        except Exception as e:
            line["strong_80m_1gpu"] = {"error": repr(e)[:300]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--points-per-gpu", type=int, default=10_000_000)
    ap.add_argument("--workload", choices=["uniform", "dambreak", "twoset"], default="uniform")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="only the headline workload: skip configs 3 / 4, the 80M single-GPU run, the micro-benchmark and the C++ arm")
    ap.add_argument("--shard-input", choices=["slab", "random"], default="slab",
                    help="multi-GPU only: every rank starts with its own Z slab of the cloud (default; a step exchanges halo + migrants) "
                         "or with an i.i.d. sample of the whole cube (every step redistributes (N-1)/N of all points)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
