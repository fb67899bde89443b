#!/usr/bin/env python
"""bench.py — scans/s of VoFOD's per-scan volumetric hot path (BASELINE.json metric) on N B200s.

A "step" is ONE LiDAR scan (128 x 2048 rays) through the whole deterministic schedule S1 (rangefinder seeds ->
filter/voxelize -> Euclidean clustering -> close/far -> point update -> raycast accumulate + apply -> classification
+ detections -> separated-background-cluster pass) on the cfg2 map (0.5 m voxels, 200 x 200 x 80 m).  Step k is scan
k of the seeded synthetic sequence; the map state carries over, warm-up steps are the first scans of the sequence.

  value : scans/s with every scan already resident in HBM (vofod_process_scan_resident)
  e2e   : scans/s through the host-buffer entry point vofod_process_scan (pinned host scan -> H2D inside the timed
          region, result record + detections D2H inside the timed region)
  N > 1 : every rank runs its own independent scan stream on its own GPU (BASELINE configs[3]); no data-path
          collective; value = sum of scans / max-over-ranks device time  ("weak" scaling)

--impl reference times the reference's CPU path (the oracle restatement: the reference itself cannot be compiled
here — no ROS/PCL/Eigen) on the host cores with the same schedule, scans and map.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 2048, 128
VOXEL = 0.5
OPAREA = (200.0, 200.0, 80.0)
WORKLOAD = "cfg2: synthetic OS0-128 scans (128x2048), 0.5 m voxels, 200x200x80 m map (401x401x161), schedule S1, default detection_params"
BYTES_PER_TRAVERSAL = 8  # 4 B read + 4 B write of the fp32 accumulator per forEachRay callback


def make_params():
    from vofod_b200 import abi
    p = abi.default_params()
    for i, (o, s) in enumerate(zip((0.0, 0.0, -1.25), OPAREA)):
        p.oparea_offset[i] = o
        p.oparea_size[i] = s
    return p


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None"""
    path = os.path.join(ROOT, "profiles", "r01_ncu_raycast_summary.json")
    try:
        return int(json.load(open(path))["traffic_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from vofod_b200 import abi, capi, multi, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    K, Wm = args.steps, args.warmup
    n_scans = K + Wm

    v = capi.Vofod(local_rank)  # raises if libvofod_cuda.so / the device is missing: no CPU fallback
    p = make_params()
    dirs = synth.sim_lut(W, H)
    if args.no_pdl:
        v.set_option(abi.OPT_PDL, 0)
    v.reset(p, VOXEL)
    v.set_sensor(W, H, dirs)
    stream = torch.cuda.ExternalStream(v.stream(), device=torch.device("cuda", local_rank))

    # every rank owns an independent stream of scans (rank r starts its trajectory r*1000 scans later)
    N = W * H
    pinned = torch.empty((n_scans, N * abi.PT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
    host_scans = pinned.numpy().view(abi.PT_DTYPE).reshape(n_scans, N)
    poses, scheds = [], []
    for k in range(n_scans):
        # weak scaling wants the same work on every GPU: each rank feeds its own context the same seeded sequence.  (With
        # --distinct-streams rank r flies another part of the trajectory after the common take-off: multi.stream_scan_index.)
        idx = multi.stream_scan_index(rank, k) if args.distinct_streams else k
        _, pose, rp, _ = synth.generate(synth.SCENE_CITY, idx, W, H, dirs, 1.0, out=host_scans[k])
        poses.append(pose)
        scheds.append(abi.schedule_s1(rp))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    dets = np.zeros(256, dtype=abi.DETECTION_DTYPE)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_leg(resident, graph=True):
        """-> (per-step device ms list, totals dict).  State is reset, so all legs do identical work."""
        v.set_option(abi.OPT_GRAPH, int(graph))
        v.reset(p, VOXEL)
        if resident:
            for k in range(n_scans):
                v.upload_scan(k, host_scans[k])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_scans)]
        tot = {"trav": 0, "ray_ms": 0.0, "dets": 0, "stage": {}}
        l0 = v.kernel_launches()
        whole = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        barrier()
        for k in range(n_scans):
            if k == Wm:
                barrier()
                l0 = v.kernel_launches()
                whole[0].record(stream)
            with torch.cuda.stream(stream):
                flush.fill_(k & 0xFF)  # L2 flush between steps, outside the timed events
                ev[k][0].record(stream)
                if resident:
                    res, d = v.process_scan_resident(k, poses[k], p, scheds[k], dets=dets)
                else:
                    # streaming sensor: the next scan's H2D copy is announced before this scan is processed, so it overlaps
                    # this scan's kernels; every copy still happens inside the timed loop
                    if k + 1 < n_scans:
                        v.prefetch_scan(host_scans[k + 1])
                    res, d = v.process_scan(host_scans[k], poses[k], p, scheds[k])
                ev[k][1].record(stream)
            if k >= Wm:
                st = v.stage_times()
                tot["trav"] += res.n_traversals
                tot["ray_ms"] += st["raycasting"]
                tot["dets"] += res.n_detections
                for name, ms in st.items():
                    tot["stage"][name] = tot["stage"].get(name, 0.0) + ms
        whole[1].record(stream)
        barrier()
        tot["launches"] = v.kernel_launches() - l0
        ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(Wm, n_scans)]
        # one bracket around all K steps: additionally contains the L2-flush writes and the host time between two scans
        tot["whole_ms"] = whole[0].elapsed_time(whole[1])
        return ms, tot

    if args.profile_leg:
        ms, tot = run_leg(True, graph=args.profile_leg == "graph")
        order = sorted(range(len(ms)), key=lambda i: -ms[i])[:6]
        print(json.dumps({"profile_leg": args.profile_leg, "ms_per_step": sum(ms) / K, "median_ms": float(np.median(ms)),
                          "slowest_steps": [(Wm + i, round(ms[i], 3)) for i in order], "replay_stats": v.stats(),
                          "stage_ms_per_step": {k: round(x / K, 4) for k, x in tot["stage"].items()}}))
        v.close()
        return
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_res, tot_res = run_leg(True)
    ms_e2e, tot_e2e = run_leg(False)
    clocks = sampler.stop()
    stats_after_graph_legs = v.stats()
    # per-stage device times (CUDA events between the stages on the library's stream) need the kernel-by-kernel path:
    # the same sequence once more with graph replay switched off.  Only the stage table and the roofline use it.
    ms_eager, tot_eager = run_leg(True, graph=False)
    tot_res["ray_ms"] = tot_eager["ray_ms"]
    tot_res["stage"] = tot_eager["stage"]
    assert tot_eager["trav"] == tot_res["trav"]

    t_res = torch.tensor([sum(ms_res), sum(ms_e2e), float(tot_res["trav"]), tot_res["ray_ms"], tot_res["whole_ms"], tot_e2e["whole_ms"]], dtype=torch.float64,
                         device="cuda")
    if world > 1:
        tmax = t_res.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_res.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax, tsum = t_res, t_res
    total_ms, total_ms_e2e = float(tmax[0]), float(tmax[1])
    trav_all, ray_ms_max = float(tsum[2]), float(tmax[3])
    whole_ms_max, whole_e2e_ms_max = float(tmax[4]), float(tmax[5])

    out = None
    if rank == 0:
        peak, peak_kind = peaks()
        # dominant kernel: k_raycast_accumulate — algorithmic bytes = traversals x 8 B, duration = the "raycasting" stage
        # (CUDA events on the context's stream around that single launch), averaged per launch over the timed steps
        trav_per_launch = tot_res["trav"] / K
        ray_ms_per_launch = tot_res["ray_ms"] / K
        achieved = trav_per_launch * BYTES_PER_TRAVERSAL / (ray_ms_per_launch * 1e-3) / 1e9 if ray_ms_per_launch > 0 else 0.0
        out = {
            "metric": "scans/s",
            "value": world * K / (total_ms * 1e-3),
            "unit": "scans/s",
            "n_gpus": world,
            "steps": K,
            "warmup": Wm,
            "ms_per_step": total_ms / K,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32 scores / u64 fixed-point path lengths / u32 keys",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "scans_per_rank": K, "rays_per_scan": N,
                       "l2": "flushed between steps (256 MB write on the same stream, outside the timed events)",
                       "parallelism": (f"{world} independent scan streams (" + ("distinct trajectories" if args.distinct_streams else "same seeded sequence on every rank")
                                       + "), one context per GPU, no collective") if world > 1 else "1 GPU"},
            "gvoxel_traversals_per_s": trav_all / (ray_ms_max * 1e-3) / 1e9 if ray_ms_max > 0 else None,
            "gvoxel_traversals_per_s_full_path": trav_all / (total_ms * 1e-3) / 1e9,
            "traversals_per_scan": trav_per_launch,
            "e2e": {"value": world * K / (total_ms_e2e * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": N * abi.PT_DTYPE.itemsize,
                    "d2h_bytes_per_step": 64 * 8, "ms_per_step": total_ms_e2e / K},  # the 64 result counters; detection records (168 B each) follow only when a scan has detections
            "single_bracket": {"note": "one event pair around all K steps of each leg: includes the 256 MB L2-flush write per step and the host time "
                                       "between two synchronous calls, which the per-step events leave out",
                               "value": world * K / (whole_ms_max * 1e-3), "e2e": world * K / (whole_e2e_ms_max * 1e-3), "unit": "scans/s"},
            "gpu_launches": int(tot_res["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_raycast_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_kind": peak_kind, "algorithmic_bytes_per_launch": trav_per_launch * BYTES_PER_TRAVERSAL,
                         "ms_per_launch": ray_ms_per_launch},
            "stage_ms_per_step": {k: round(x / K, 4) for k, x in tot_res["stage"].items()},
            "stage_note": "stage times and the roofline kernel time come from a third, kernel-by-kernel leg (graph replay off): %.4f ms/step" % (sum(ms_eager) / K),
            "detections_in_timed_steps": int(tot_res["dets"]),
            "replay_stats": stats_after_graph_legs,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(min(12, n_scans))
    del stream, flush
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    v.close()
    if not args.no_slab:
        # BASELINE.json configs[4]: the 6.4 GB map cut into `world` slabs, strong scaling (the same scans whatever N), both ray lengths
        rec = slab_record(local_rank, rank, world, args.slab_steps, args.slab_warmup)
        if out is not None:
            out["slab_cfg5"] = rec
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


SLAB_WORKLOAD = ("cfg5 large map: 0.25 m voxels, 500x500x100 m (2001x2001x401 cells = 6.4 GB fp32 grid), cut into x-slabs (one per GPU) + 16-cell halo, "
                 "whole schedule S1 (seeds, filter/voxelize, cluster, close/far, point update, clipped raycast + apply, classification + detections on exchanged "
                 "map boxes, separated-background pass on the gathered voxel lists)")


def slab_leg(local_rank, rank, world, raycast_max, K, Wm, scans=None):
    """scans/s of the slab path on `world` slabs (device-timed on the library's stream, max over ranks).  The scan of step k starts in
    pinned host memory on rank 0: H2D + NCCL broadcast are inside the timed region."""
    import torch
    import torch.distributed as dist

    from vofod_b200 import abi, capi, slab, synth
    p = abi.default_params()
    for i, (o, sz) in enumerate(zip((0.0, 0.0, -1.25), (500.0, 500.0, 100.0))):
        p.oparea_offset[i] = o
        p.oparea_size[i] = sz
    p.raycast_max_distance = float(raycast_max)
    dirs = synth.sim_lut(W, H)
    v = capi.Vofod(local_rank)
    worker = slab.SlabWorker(v, p, 0.25, (W, H), dirs, rank, world, halo=16)
    N = W * H
    n_scans = K + Wm
    if scans is None:
        pinned = torch.empty((n_scans, N * abi.PT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
        host = pinned.numpy().view(abi.PT_DTYPE).reshape(n_scans, N)
        poses, scheds = [], []
        for k in range(n_scans):
            # every rank generates the (deterministic) pose and seed point; only rank 0's scan buffer is ever read
            _, pose, rp, _ = synth.generate(synth.SCENE_CITY, k, W, H, dirs, 2.5, out=host[k])
            poses.append(pose)
            scheds.append(abi.schedule_s1(rp))
        scans = (pinned, host, poses, scheds)
    pinned, host, poses, scheds = scans
    stream = torch.cuda.ExternalStream(v.stream(), device=torch.device("cuda", local_rank))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_scans)]
    trav, dets, l0 = 0, 0, 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    for k in range(n_scans):
        if k == Wm:
            barrier()
            l0 = v.kernel_launches()
        ev[k][0].record(stream)
        res, d = worker.step(host[k], poses[k], scheds[k])
        ev[k][1].record(stream)
        if k >= Wm:
            trav += res.n_traversals  # summed over the slabs by the library (a traversal is counted by the slab that owns the voxel)
            dets += res.n_detections
    barrier()
    launches = v.kernel_launches() - l0
    ms = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(Wm, n_scans))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    mi = v.map_info()
    cells = int(mi.storage_size[0]) * int(mi.storage_size[1]) * int(mi.storage_size[2])
    del stream
    torch.cuda.synchronize()
    worker.close()
    torch.cuda.empty_cache()
    return {"value": K / (total_ms * 1e-3), "unit": "scans/s", "ms_per_step": total_ms / K, "n_slabs": world, "steps": K, "warmup": Wm,
            "raycast_max_distance_m": raycast_max, "traversals_per_scan": trav / K, "gvoxel_traversals_per_s_full_path": trav / (total_ms * 1e-3) / 1e9,
            "detections_in_timed_steps": dets, "gpu_launches_rank0": int(launches), "slab0_storage_cells": cells,
            "h2d_bytes_per_step": N * abi.PT_DTYPE.itemsize, "scaling": "strong"}, scans


def slab_record(local_rank, rank, world, K, Wm, dists=(20.0, 200.0)):
    """cfg5 sub-record: scans/s on `world` slabs for both ray lengths and, when world > 1, the same scans on ONE GPU (rank 0, the other
    ranks wait) measured in the same run, so that the strong-scaling efficiency needs no second file"""
    import torch.distributed as dist
    rec = {"workload": SLAB_WORKLOAD, "parallelism": f"slab{world}: ncclBroadcast of the packed scan (5.2 MB) + ncclAllReduce / ncclAllGather of the exchange buffers "
                                                     "inside libvofod_cuda, one host wait per scan" if world > 1 else "1 slab = the whole grid on one GPU"}
    for d in dists:
        r, scans = slab_leg(local_rank, rank, world, d, K, Wm)
        key = "%gm" % d
        rec[key] = r
        if world > 1:
            if rank == 0:
                one, _ = slab_leg(local_rank, 0, 1, d, K, Wm, scans=scans)
                r["one_gpu_same_run"] = {"value": one["value"], "ms_per_step": one["ms_per_step"]}
                r["strong_scaling_efficiency"] = r["value"] / one["value"] / world
            dist.barrier()
    return rec


def run_slab(args):
    """Large-map mode (BASELINE.json configs[4]) on its own: `--mode slab [--raycast-max D]`.  Strong scaling: the same scans on 1..N GPUs."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    rec = slab_record(local_rank, rank, world, args.steps, args.warmup, dists=(float(args.raycast_max),))
    r = rec["%gm" % float(args.raycast_max)]
    if rank == 0:
        print(json.dumps({
            "metric": "scans/s", "value": r["value"], "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 scores / u64 fixed-point path lengths", "data": "synthetic",
            "config": {"workload": SLAB_WORKLOAD + ", raycast.max_distance %g m" % args.raycast_max, "parallelism": rec["parallelism"],
                       "l2": "grids (GBs) far larger than L2"}, "mode": "slab", "slab": r}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def cpu_baseline(n_scans, timed_from=2):
    """The oracle (CPU restatement of the reference's path) on the host, single thread, first scans of the same sequence."""
    from oracle import oracle  # the ONLY use of oracle/ in this file besides --impl reference: the reported CPU baseline
    from vofod_b200 import abi, synth
    p = make_params()
    dirs = synth.sim_lut(W, H)
    o = oracle.Oracle(track_counts=False)
    o.reset(p, VOXEL)
    o.set_sensor(W, H, dirs)
    t_total, n = 0.0, 0
    for k in range(n_scans):
        scan, pose, rp, _ = synth.generate(synth.SCENE_CITY, k, W, H, dirs)
        s = abi.schedule_s1(rp)
        t0 = time.perf_counter()
        o.process_scan(scan, pose, p, s)
        dt = time.perf_counter() - t0
        if k >= timed_from:
            t_total += dt
            n += 1
    o.close()
    return {"value": n / t_total if t_total > 0 else None, "unit": "scans/s", "cores": 1, "kind": "port",
            "sample": f"scans {timed_from}..{n_scans - 1} of the same sequence, schedule S1, g++ -O3 -DNDEBUG (reference flags), 1 thread; host has {os.cpu_count()} cpus"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from vofod_b200 import abi, synth
    K, Wm = args.steps, args.warmup
    if K + Wm > 120:  # bounded sample: ~1.5 s of CPU work per scan
        K, Wm = min(K, 100), min(Wm, 20)
    p = make_params()
    dirs = synth.sim_lut(W, H)
    o = oracle.Oracle(track_counts=False)
    o.reset(p, VOXEL)
    o.set_sensor(W, H, dirs)
    t_total, trav, t_ray = 0.0, 0, 0.0
    for k in range(K + Wm):
        scan, pose, rp, _ = synth.generate(synth.SCENE_CITY, k, W, H, dirs)
        s = abi.schedule_s1(rp)
        t0 = time.perf_counter()
        res, _ = o.process_scan(scan, pose, p, s)
        dt = time.perf_counter() - t0
        if k >= Wm:
            t_total += dt
            trav += res.n_traversals
            t_ray += o.stage_times()[5] * 1e-3
    o.close()
    val = K / t_total
    sample = (f"{K} scans (after {Wm} warm-up scans) of the same sequence, schedule S1 fully serial on 1 thread — the reference runs each of its actors "
              f"(scan, raycast, background clusters) on a single thread (pointcloud_threads: 1); host has {os.cpu_count()} cpus")
    out = {"impl": "reference", "metric": "scans/s", "value": val, "unit": "scans/s", "n_gpus": args.gpus, "steps": K, "warmup": Wm,
           "ms_per_step": 1e3 * t_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": "CPU oracle (restatement of the reference; the reference needs ROS/PCL/Eigen and cannot be built here)"},
           "gvoxel_traversals_per_s": trav / t_ray / 1e9 if t_ray > 0 else None,
           "cpu_baseline": {"value": val, "unit": "scans/s", "cores": 1, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=80)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pdl", action="store_true", help="A/B switch: launch the kernels without programmatic dependent launch")
    ap.add_argument("--distinct-streams", action="store_true", help="N > 1: every rank processes a different part of the trajectory instead of the same sequence")
    ap.add_argument("--mode", default="streams", choices=["streams", "slab"],
                    help="streams (default, the driver's contract): cfg2, one independent scan stream per GPU; slab: cfg5 large map cut into x-slabs")
    ap.add_argument("--no-slab", action="store_true", help="skip the cfg5 slab sub-record")
    ap.add_argument("--slab-steps", type=int, default=20)
    ap.add_argument("--slab-warmup", type=int, default=6)
    ap.add_argument("--raycast-max", type=float, default=20.0, help="slab mode: raycast.max_distance [m] (yaml default 20, dynamic_reconfigure maximum 200)")
    ap.add_argument("--profile-leg", default="", choices=["", "graph", "eager"],
                    help="profiling aid: run ONLY the HBM-resident leg (graph replay or kernel-by-kernel) and print nothing the driver parses")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "slab":
        run_slab(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
