#!/usr/bin/env python
"""bench.py -- million neighbour-queries/s (build + query) of the B200 engine, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points-per-gpu P] [--workload uniform|dambreak]

A step is one full pass of the hot path (everything tns::TreeNSearch::run() does: world box, Morton keys, radix sort, reorder,
cell start/end, 27-cell query, neighbour lists) over one synthetic cloud.

  N = 1 : BASELINE.json configs[1]: 10M uniform-random points, single set, fixed radius (k_mean ~ 29.7).
  N > 1 : the same density with 10M points PER GPU (weak scaling; N = 8 is configs[4], 80M points), Z-slab sharded with a
          one-cell halo exchanged over NCCL (treensearch_b200/sharded.py).

`value`  : inputs resident in HBM, lists left in HBM, timed with CUDA events on the stream the kernels run on.
`e2e`    : the same metric through the public API with HOST buffers: pinned host points in, host-addressable lists out.
`--impl reference` : the unmodified reference's run() (oracle/_ref, AVX2 + OpenMP, all host threads) on the same workload.
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
CPU_SAMPLE_POINTS = 10_000_000          # the reference arm / cpu_baseline always runs at most this many points


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


# ------------------------------------------------------------------------------------------------------------ reference arm
def time_reference(pts, r, steps, warmup):
    """The unmodified reference on the host cores: zsort once (its intended regime, README.md:142-143), then time run()."""
    from oracle import loader
    ref = loader.Reference()
    pts = np.ascontiguousarray(pts).copy()
    ref.set_search_radius(r)
    ref.add_point_set(pts)
    ref.set_active_search(0, 0, True)
    cores = loader.Reference.n_threads()
    # raw (unsorted) input: one warm-up + one timed run
    ref.time_runs(1)
    ms_unsorted = float(ref.time_runs(1)[0])
    ref.prepare_zsort()
    ref.apply_zsort(0, pts.reshape(-1), 3)
    ref.time_runs(max(warmup, 1))
    ms = ref.time_runs(steps)
    n = pts.shape[0]
    return {
        "ms_per_step": float(np.mean(ms)), "ms_best": float(np.min(ms)), "ms_unsorted_input": ms_unsorted,
        "value": n / (float(np.mean(ms)) * 1e-3) / 1e6, "value_unsorted_input": n / (ms_unsorted * 1e-3) / 1e6,
        "cores": cores, "n_points": n,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import loader
    if not loader.reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtns_ref.so missing (make -C oracle where /root/reference exists)"}))
        return
    total = args.points_per_gpu * args.gpus
    n = min(total, CPU_SAMPLE_POINTS)
    pts, _ = make_cloud(args.workload, n)
    # same density per search sphere as the GPU arm's cloud
    from treensearch_b200 import clouds
    r = float(clouds.radius_for_mean_neighbors(n)) if args.workload == "uniform" else make_cloud(args.workload, n)[1]
    t = time_reference(pts, r, args.steps, args.warmup)
    sample = f"{n} of {total} points ({args.workload}, same k_mean), z-sorted input, {args.steps} run() calls after {max(args.warmup, 1)} warm-up"
    line = {
        "metric": METRIC, "value": t["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(args, total),
        "cpu_baseline": {"value": t["value"], "unit": UNIT, "cores": t["cores"], "kind": "reference", "sample": sample,
                         "value_unsorted_input": t["value_unsorted_input"]},
        "e2e": {"value": t["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, total):
    name = {"uniform": "uniform-random points in the unit cube, single set, fixed radius (k_mean~29.7)",
            "dambreak": "SPH dam-break clustered points (~60 neighbours interior), single set, fixed radius"}[args.workload]
    return {"workload": f"{total} {name}" + (f", Z-slab sharded over {args.gpus} GPUs with NCCL halo exchange" if args.gpus > 1 else ", 1xB200"),
            "n_points": total, "points_per_gpu": args.points_per_gpu, "active_searches": "0->0",
            **({"shard_input": {"slab": "every rank holds its own Z slab (halo + migrating points exchanged per step)",
                                "random": "every rank holds an i.i.d. sample of the whole cube (full redistribution per step)"}[args.shard_input]}
               if args.gpus > 1 else {}),
            "l2": "flushed between timed steps (256 MiB write, untimed); per-step working set ~1.9 GB >> 126 MB L2"}


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
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")

    total = args.points_per_gpu * args.gpus
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        pts_np, r = make_cloud(args.workload, total)
        n_local = total
        # ---- device-resident arm
        eng = t.TreeNSearch(local_rank)
        eng.set_stream(stream.cuda_stream)
        eng.set_option(t.TNSB_OPT_HOST_RESULTS, 0)
        d_pts = torch.from_numpy(pts_np).cuda()
        eng.set_search_radius(r)
        eng.add_point_set(d_pts)
        eng.set_active_search(0, 0, True)
        step_dev = eng.run
        # ---- end-to-end arm: pinned host points in, host-addressable lists out (the call a user of the reference makes)
        eng_e2e = t.TreeNSearch(local_rank)
        eng_e2e.set_stream(stream.cuda_stream)
        h_pts = torch.from_numpy(pts_np).pin_memory()
        eng_e2e.set_search_radius(r)
        eng_e2e.add_point_set(h_pts)
        eng_e2e.set_active_search(0, 0, True)
        step_e2e = eng_e2e.run
        stats_of = eng.stats
        stats_e2e = eng_e2e.stats
    else:
        from treensearch_b200 import sharded
        job = sharded.ShardedUniformJob(args.workload, args.points_per_gpu, rank, world, local_rank, stream, args.shard_input)
        r = job.radius
        n_local = args.points_per_gpu
        step_dev = job.step_device
        step_e2e = job.step_e2e
        stats_of = job.stats
        stats_e2e = job.stats_e2e

    def timed(step, k, w):
        for _ in range(w):
            step()
        barrier()
        ms = []
        for _ in range(k):
            flush.fill_(1)                      # L2 flush, outside the timed pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            step()
            e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        barrier()
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)      # max over ranks
        return float(tot.item()) / k, ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, ms_list = timed(step_dev, args.steps, max(args.warmup, 3))
    st = stats_of()
    ms_e2e, _ = timed(step_e2e, max(1, min(args.steps, 5)), 3)
    st2 = stats_e2e()
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (device-resident and end-to-end)

    # ---- roofline of the dominant kernel (the 27-cell query): algorithmic bytes per launch / CUDA-event duration
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
    e2e_value = total / (ms_e2e * 1e-3) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, total),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(st2["h2d_bytes"]) * world, "d2h_bytes_per_step": int(st2["d2h_bytes"]) * world,
                "note": "pinned host points in; count-prefixed neighbour lists written by the query kernel into mapped pinned host memory "
                        "(zero-copy, PCIe-bound) + list_pos table copied back" + ("; bytes = rank 0 x n_gpus" if world > 1 else "")},
        "gpu_launches": int(st["n_kernel_launches"]) * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "query_kernel (27-cell query)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms": q_ms, "algorithmic_bytes_per_launch": q_bytes / max(st["n_query_launches"], 1),
                     "whole_run": {"algorithmic_bytes": run_bytes, "ms": st["ms_total_device"],
                                   "achieved": run_bytes / (st["ms_total_device"] * 1e-3) / 1e9 if st["ms_total_device"] > 0 else 0.0,
                                   "frac": run_bytes / (st["ms_total_device"] * 1e-3) / 1e9 / peak if st["ms_total_device"] > 0 else 0.0}},
        "stages_ms": {k: st[k] for k in ("ms_aabb", "ms_keys", "ms_sort", "ms_reorder", "ms_cells", "ms_query", "ms_total_device")},
        "e2e_stages_ms": {k: st2[k] for k in ("ms_upload", "ms_total_device", "ms_download", "ms_wall")},
        "neighbors_per_query": st["n_neighbors"] / max(st["n_queries"], 1),
        "step_ms_all": ms_list,
    }

    # ---- CPU baseline beside it (rank 0, N = 1 only): the unmodified reference on a bounded sample of the same workload
    if world == 1 and not args.no_cpu_baseline:
        from oracle import loader
        if loader.reference_available():
            n = min(total, CPU_SAMPLE_POINTS)
            pts_cpu, r_cpu = (pts_np, r) if n == total else make_cloud(args.workload, n)
            tr = time_reference(pts_cpu, r_cpu, 3, 1)
            line["cpu_baseline"] = {"value": tr["value"], "unit": UNIT, "cores": tr["cores"], "kind": "reference",
                                    "sample": f"{n} points of the same workload, z-sorted input, mean of 3 run() after 1 warm-up",
                                    "ms_per_step": tr["ms_per_step"], "value_unsorted_input": tr["value_unsorted_input"]}
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
            line["cpu_baseline"] = {"value": n / dt / 1e6, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{n} uniform points, restated grid oracle (oracle/_ref not present)"}
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
    ap.add_argument("--workload", choices=["uniform", "dambreak"], default="uniform")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
