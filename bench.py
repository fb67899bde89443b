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
    for name in ("r02_ncu_raycast_summary.json", "r01_ncu_raycast_summary.json"):
        try:
            return int(json.load(open(os.path.join(ROOT, "profiles", name)))["traffic_bytes_per_launch"])
        except Exception:
            continue
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def pin_to_gpu_numa_node(local_rank):
    """Run this rank's host side (and place its pinned buffers, first touch) on the NUMA node its GPU hangs off: N processes that all
    pull their scans through one node's memory controllers is what capped the 8-GPU e2e scaling in round 1."""
    info = {"applied": False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info.update({"pci": bdf, "node": node})
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info.update({"applied": True, "n_cpus": len(cpus)})
    except Exception as e:  # not fatal: unpinned is what round 1 measured
        info["error"] = str(e)[:80]
    return info


def l2_peaks():
    """L2 yardsticks measured on this pool's B200 by tools/l2_microbench.cu (SURVEY.md §8d asks for them next to the HBM figure)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_l2_microbench.json")))
    except Exception:
        return None


def whole_scan_bytes(n_rays, m_vox, traversals, window_cells, grid_cells, n_bg_cells):
    """SURVEY.md §8d byte table for one scan of schedule S1.  Two totals: `algorithmic_min` counts the grid passes over the cells they can
    concern (ray window, raised chunks), `reference_faithful` counts them over the whole grid as the reference executes them."""
    ingest = n_rays * (20 + 24 + 1)
    trav = traversals * BYTES_PER_TRAVERSAL
    vg = n_rays * (12 + 8) + 4 * n_rays * 16 + m_vox * 16
    upd = m_vox * (16 + 8 + 4)
    clus = m_vox * (12 + 8 + 27 * 4 + 4 * 4)
    apply_win = window_cells * 28
    apply_full = grid_cells * 28
    nover_full = grid_cells * 4
    sep_full = grid_cells * 12
    sep_min = n_bg_cells * (12 + 16)
    common = ingest + trav + vg + upd + clus
    return {"algorithmic_min": common + apply_win + n_bg_cells * 4 + sep_min, "reference_faithful": common + apply_full + nover_full + sep_full,
            "terms": {"scan_ingest": ingest, "traversal": trav, "voxel_grid": vg, "point_update": upd, "clustering": clus, "ray_apply_windowed": apply_win,
                      "ray_apply_full_grid": apply_full, "n_voxels_over_full_grid": nover_full, "sepclusters_full_grid": sep_full}}


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
    numa = pin_to_gpu_numa_node(local_rank) if not args.no_numa_pin else {"applied": False, "off": True}
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
        sch = abi.schedule_s1(rp)
        # the background thread's pass of scan k is carried out at the start of step k + 1, beside that scan's front end
        # (vofod_schedule::sep_deferred): every step still contains exactly one pass, the order of all map operations is S1's
        sch.sep_deferred = 0 if args.no_defer_sep else 1
        scheds.append(sch)
    scheds_inline = []
    for sch in scheds:
        c = abi.Schedule.from_buffer_copy(sch)
        c.sep_deferred = 0
        scheds_inline.append(c)

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
        tot = {"trav": 0, "ray_ms": 0.0, "dets": 0, "stage": {}, "m_vox": 0, "n_bg": 0}
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
                sk = scheds[k] if graph else scheds_inline[k]  # the kernel-by-kernel leg feeds the stage table: pass in line, in its own stage
                if resident:
                    res, d = v.process_scan_resident(k, poses[k], p, sk, dets=dets)
                else:
                    # streaming sensor: the next scan's H2D copy is announced before this scan is processed, so it overlaps
                    # this scan's kernels; every copy still happens inside the timed loop
                    if k + 1 < n_scans:
                        v.prefetch_scan(host_scans[k + 1])
                    res, d = v.process_scan(host_scans[k], poses[k], p, sk)
                ev[k][1].record(stream)
            if k >= Wm:
                st = v.stage_times()
                tot["trav"] += res.n_traversals
                tot["ray_ms"] += st["raycasting"]
                tot["dets"] += res.n_detections
                tot["m_vox"] += res.n_voxels
                tot["n_bg"] += res.n_bg
                for name, ms in st.items():
                    tot["stage"][name] = tot["stage"].get(name, 0.0) + ms
        whole[1].record(stream)
        barrier()
        tot["launches"] = v.kernel_launches() - l0
        ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(Wm, n_scans)]
        # one bracket around all K steps: additionally contains the L2-flush writes and the host time between two scans
        tot["whole_ms"] = whole[0].elapsed_time(whole[1])
        return ms, tot

    def run_streaming():
        """e2e as a user gets it: K scans from pinned host buffers through ONE call of the C ABI (vofod_process_scan_batch: scan k + 1's H2D
        copy overlaps scan k's kernels, results read back per scan), ONE event bracket around all K scans, no L2 flush — every step's input
        is a new 5.2 MB scan from host memory (K x 5.2 MB over the leg), the map stays wherever the previous scan left it, as in production."""
        v.set_option(abi.OPT_GRAPH, 1)
        v.reset(p, VOXEL)
        v.process_scan_batch([host_scans[k] for k in range(Wm)], poses[:Wm], p, scheds[:Wm])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = v.kernel_launches()
        with torch.cuda.stream(stream):
            e0.record(stream)
            res, done = v.process_scan_batch([host_scans[k] for k in range(Wm, n_scans)], poses[Wm:], p, scheds[Wm:])
            v.flush()  # the last scan's deferred pass belongs to the timed region
            e1.record(stream)
        barrier()
        assert done == K
        return e0.elapsed_time(e1), sum(r.n_traversals for r in res), sum(r.n_detections for r in res), v.kernel_launches() - l0

    def run_streams_on_this_gpu(S=8):
        """BASELINE.json configs[3] on the GPUs at hand: S independent scan streams (S contexts, S maps, S host threads) share this GPU; every
        stream runs the e2e leg (vofod_process_scan_batch from pinned host buffers).  A single scan is a chain of short dependent kernels that
        leaves most of the machine idle; independent streams fill it.  Timed on the host (start barrier -> last stream done + device idle)."""
        import threading
        import time
        ctxs = [capi.Vofod(local_rank) for _ in range(S)]
        for c in ctxs:
            c.reset(p, VOXEL)
            c.set_sensor(W, H, dirs)
            c.process_scan_batch([host_scans[k] for k in range(Wm)], poses[:Wm], p, scheds[:Wm])
        torch.cuda.synchronize()
        start = threading.Barrier(S + 1)
        done_n = [0] * S

        def work(i):
            start.wait()
            res, done = ctxs[i].process_scan_batch([host_scans[k] for k in range(Wm, n_scans)], poses[Wm:], p, scheds[Wm:])
            ctxs[i].flush()
            done_n[i] = done
        th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
        for t in th:
            t.start()
        start.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert all(d == K for d in done_n), done_n
        for c in ctxs:
            c.close()
        return {"workload": f"{S} independent scan streams (cfg2 scans, one context and one map each) on ONE GPU, each through vofod_process_scan_batch from pinned "
                            "host buffers; host-timed from a common start to the last stream's last result",
                "streams": S, "value": S * K / dt, "unit": "scans/s", "ms_per_scan": dt * 1e3 / (S * K), "h2d_bytes_per_step": N * abi.PT_DTYPE.itemsize}

    def run_cfg3(K3=40, W3=32, scene=None, what=None):
        """BASELINE.json configs[2]: Gazebo-like scene with 3 sphere UAVs — the scans that DO produce detections (exploreToGround, frontier
        write-back, submap confidence) — same map, same per-step timing as the resident leg.  scene = SCENE_SWARM: the same with 200 spheres
        (~230 far clusters per scan through the one-block sequential classification kernel)"""
        scene = synth.SCENE_GAZEBO if scene is None else scene
        v.set_option(abi.OPT_GRAPH, 1)
        v.reset(p, VOXEL)
        n3 = K3 + W3
        buf = np.zeros((n3, N), dtype=abi.PT_DTYPE)
        ps, ss = [], []
        for k in range(n3):
            _, pose, rp, _ = synth.generate(scene, k, W, H, dirs, 1.0, out=buf[k])
            ps.append(pose)
            ss.append(abi.schedule_s1(rp))
            v.upload_scan(k, buf[k])
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n3)]
        nd = nfar = 0
        torch.cuda.synchronize()  # (rank 0 runs this leg alone: no collective in here — a dist.barrier would pair up with the other ranks' NEXT one)
        for k in range(n3):
            with torch.cuda.stream(stream):
                flush.fill_(k & 0xFF)
                ev3[k][0].record(stream)
                res, d = v.process_scan_resident(k, ps[k], p, ss[k], dets=dets)
                ev3[k][1].record(stream)
            if k >= W3:
                nd += res.n_detections
                nfar += res.n_far_clusters
        torch.cuda.synchronize()
        ms3 = sum(ev3[k][0].elapsed_time(ev3[k][1]) for k in range(W3, n3))
        return {"workload": what or "cfg3: Gazebo-like scene (ground + 4 buildings + 3 sphere UAVs), 128x2048 rays, cfg2 map, schedule S1, scans resident in HBM",
                "value": K3 / (ms3 * 1e-3), "unit": "scans/s", "ms_per_step": ms3 / K3, "steps": K3, "warmup": W3, "detections_in_timed_steps": int(nd),
                "far_clusters_per_scan": nfar / K3}

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
    stream_ms, stream_trav, stream_dets, stream_launches = run_streaming()
    clocks = sampler.stop()
    def sub_record(fn, *a, **kw):
        """a sub-record never takes the headline line down with it (these legs run on one rank: no collective inside, see tests/test_host_logic.py)"""
        try:
            return fn(*a, **kw)
        except Exception as e:  # noqa: BLE001
            return {"error": "%s: %s" % (type(e).__name__, e)}
    streams8 = sub_record(run_streams_on_this_gpu, 8) if world == 1 and not args.no_cfg3 else None
    cfg3 = sub_record(run_cfg3) if rank == 0 and not args.no_cfg3 else None
    swarm = sub_record(run_cfg3, scene=synth.SCENE_SWARM, what="stress: the cfg3 scene with 200 sphere UAVs on rings around the sensor (~230 far clusters, > 100 "
                       "detections per scan), 128x2048 rays, cfg2 map, schedule S1, scans resident in HBM") if rank == 0 and not args.no_cfg3 else None
    if world > 1:
        dist.barrier()
    stats_after_graph_legs = v.stats()
    # per-stage device times (CUDA events between the stages on the library's stream) need the kernel-by-kernel path:
    # the same sequence once more with graph replay switched off.  Only the stage table and the roofline use it.
    ms_eager, tot_eager = run_leg(True, graph=False)
    tot_res["ray_ms"] = tot_eager["ray_ms"]
    tot_res["stage"] = tot_eager["stage"]
    assert tot_eager["trav"] == tot_res["trav"]

    t_res = torch.tensor([sum(ms_res), sum(ms_e2e), float(tot_res["trav"]), tot_res["ray_ms"], tot_res["whole_ms"], tot_e2e["whole_ms"], stream_ms], dtype=torch.float64,
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
    stream_ms_max = float(tmax[6])

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
                       "schedule": "S1" + ("" if args.no_defer_sep else "; the separated-background pass of scan k runs at the start of step k+1 beside that scan's "
                                          "front end (sep_deferred): one pass per step, same order of map operations"),
                       "parallelism": (f"{world} independent scan streams (" + ("distinct trajectories" if args.distinct_streams else "same seeded sequence on every rank")
                                       + "), one context per GPU, no collective") if world > 1 else "1 GPU"},
            "gvoxel_traversals_per_s": trav_all / (ray_ms_max * 1e-3) / 1e9 if ray_ms_max > 0 else None,
            "gvoxel_traversals_per_s_full_path": trav_all / (total_ms * 1e-3) / 1e9,
            "traversals_per_scan": trav_per_launch,
            # headline: the streaming leg — ONE event bracket around K scans fed from pinned host memory through one C-ABI call, no flush
            "e2e": {"value": world * K / (stream_ms_max * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": N * abi.PT_DTYPE.itemsize,
                    "d2h_bytes_per_step": 64 * 8, "ms_per_step": stream_ms_max / K,  # the 64 result counters; detection records (168 B each) follow only when a scan has detections
                    "how": "vofod_process_scan_batch: K scans from pinned host buffers, scan k+1's H2D copy overlapped with scan k's kernels, results read back per "
                           "scan; one CUDA-event bracket on the library's stream around the whole call; no L2 flush (every step's input is a new scan from host "
                           "memory; the map stays where the previous scan left it)", "gpu_launches": int(stream_launches)},
            "e2e_flushed_per_step": {"value": world * K / (total_ms_e2e * 1e-3), "unit": "scans/s", "ms_per_step": total_ms_e2e / K,
                                     "note": "round-1 definition: one vofod_process_scan call per step from Python, per-step events, 256 MB L2 flush before every step"},
            "single_bracket": {"note": "one event pair around all K steps of the two per-step legs: includes the 256 MB L2-flush write per step (~40 us) and the host "
                                       "time between two synchronous calls from Python, which the per-step events leave out",
                               "value": world * K / (whole_ms_max * 1e-3), "e2e": world * K / (whole_e2e_ms_max * 1e-3), "unit": "scans/s"},
            "numa": numa,
            "gpu_launches": int(tot_res["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_raycast_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_kind": peak_kind, "algorithmic_bytes_per_launch": trav_per_launch * BYTES_PER_TRAVERSAL,
                         "ms_per_launch": ray_ms_per_launch},
            "stage_ms_per_step": {k: round(x / K, 4) for k, x in tot_res["stage"].items()},
            "stage_note": "stage times and the roofline kernel time come from a third, kernel-by-kernel leg (graph replay off): %.4f ms/step" % (sum(ms_eager) / K),
            "detections_in_timed_steps": int(tot_res["dets"]),
            "cfg3": cfg3,
            "cfg4_streams_on_one_gpu": streams8,
            "cfg3_swarm_200_uavs": swarm,
            "replay_stats": stats_after_graph_legs,
            "clocks": clocks,
        }
        l2 = l2_peaks()
        if l2:
            # the memory-side yardsticks of an L2-resident window (tools/l2_microbench.cu): RED lane-operations per second against the
            # measured random-address RED rate, and the 8 B/traversal figure against the L2 streaming bandwidth
            reds_per_scan = l2.get("raycast_red_lane_ops_per_scan")
            out["roofline"]["l2"] = {"red_peak_gred_s": l2["red_u64_random_lane_Gred_s"], "l2_read_peak_gbs": l2["read_64MB_GBs"],
                                     "frac_of_l2_read_bw": achieved / l2["read_64MB_GBs"],
                                     "achieved_gred_s": (reds_per_scan / (ray_ms_per_launch * 1e-3) / 1e9) if reds_per_scan else None,
                                     "frac_of_red_peak": (reds_per_scan / (ray_ms_per_launch * 1e-3) / 1e9 / l2["red_u64_random_lane_Gred_s"]) if reds_per_scan else None,
                                     "floor_note": l2.get("raycast_floor_note")}
        wb = whole_scan_bytes(N, tot_res["m_vox"] / K, trav_per_launch, 87 ** 3, 401 * 401 * 161, tot_res["n_bg"] / K)
        step_s = total_ms / K * 1e-3
        out["roofline_whole_scan"] = {"bound": "hbm", "unit": "GB/s", "peak": peak, "ms_per_step": total_ms / K,
                                      "achieved_algorithmic_min": wb["algorithmic_min"] / step_s / 1e9, "frac_algorithmic_min": wb["algorithmic_min"] / step_s / 1e9 / peak,
                                      "achieved_reference_faithful": wb["reference_faithful"] / step_s / 1e9,
                                      "frac_reference_faithful": wb["reference_faithful"] / step_s / 1e9 / peak, "bytes": wb,
                                      "note": "SURVEY.md §8d byte table per scan / ms_per_step of the resident leg, per GPU; the reference-faithful total counts the "
                                              "full-grid passes (apply 28 B/cell, nVoxelsOver 4 B/cell, sepclusters 12 B/cell) that the GPU path replaces by passes over "
                                              "the ray window and the raised chunks"}
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(min(14, n_scans))
    del stream, flush
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    v.close()
    if not args.no_slab:
        # BASELINE.json configs[4]: the 6.4 GB map cut into `world` slabs, strong scaling (the same scans whatever N), both ray lengths.
        # The headline line must not depend on this sub-record: if it does not finish within its limit (a hung collective, a box short of
        # memory), every rank gives up on it, rank 0 prints the line with the reason in its place, and all ranks leave with status 0.
        import threading
        limit = float(os.environ.get("VOFOD_BENCH_SLAB_LIMIT_S", "420"))
        printed = threading.Lock()

        def give_up():
            if printed.acquire(blocking=False):
                if out is not None:
                    out["slab_cfg5"] = {"workload": SLAB_WORKLOAD, "error": "the sub-record did not finish within %g s and was abandoned" % limit}
                    print(json.dumps(out), flush=True)
                sys.stdout.flush()
                os._exit(0)
        guard = threading.Timer(limit, give_up)
        guard.daemon = True
        guard.start()
        rec = slab_record(local_rank, rank, world, args.slab_steps, args.slab_warmup)
        guard.cancel()
        if not printed.acquire(blocking=False):
            time.sleep(3600)  # (the guard is printing: it ends the process)
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


def slab_leg(local_rank, rank, world, raycast_max, K, Wm, scans=None, balanced=False):
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
    cuts = None
    if balanced and world > 1:
        # slab widths tuned to the flight area (a deployment knows where it flies): the sensor's x over the timed scans is around
        # 30 sin(0.02 k) m, i.e. cell (x + 250) / 0.25; rays reach raycast_max around it
        from vofod_b200 import multi
        ks = np.arange(Wm, Wm + K)
        centre = (float(np.mean(30.0 * np.sin(0.02 * ks))) + 250.0) / 0.25
        cuts = multi.partition_by_ray_load(2001, world, centre, raycast_max / 0.25 + 120.0)
    worker = slab.SlabWorker(v, p, 0.25, (W, H), dirs, rank, world, halo=16, cuts=cuts)
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
    parts = np.zeros(8)

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
        res, d = worker.step(host[k], poses[k], scheds[k], next_scan_host=host[k + 1] if k + 1 < n_scans else None)
        ev[k][1].record(stream)
        if k >= Wm:
            trav += res.n_traversals  # summed over the slabs by the library (a traversal is counted by the slab that owns the voxel)
            dets += res.n_detections
            parts += v.slab_times()
    barrier()
    launches = v.kernel_launches() - l0
    # where a scan's time goes on every rank: [broadcast, phase 0, exchange 0, phase 1, exchange 1, phase 2, exchange 2, phase 3] (ms, mean of the
    # timed scans; an exchange includes the wait for the slowest slab)
    pt = torch.tensor(parts / K, dtype=torch.float64, device="cuda")
    if world > 1:
        allp = [torch.zeros_like(pt) for _ in range(world)]
        dist.all_gather(allp, pt)
    else:
        allp = [pt]
    parts_by_rank = [[round(float(x), 4) for x in t.tolist()] for t in allp]
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
            "h2d_bytes_per_step": N * abi.PT_DTYPE.itemsize, "scaling": "strong",
            "slab_cut": "by ray load (multi.partition_by_ray_load)" if cuts is not None else "equal width", "own_range_rank0": [int(worker.lo), int(worker.hi)],
            "ms_by_rank_bcast_p0_x0_p1_x1_p2_x2_p3": parts_by_rank}, scans


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
            rb, _ = slab_leg(local_rank, rank, world, d, K, Wm, scans=scans, balanced=True)
            r["cut_by_ray_load"] = {k: rb[k] for k in ("value", "ms_per_step", "slab_cut", "own_range_rank0", "slab0_storage_cells", "ms_by_rank_bcast_p0_x0_p1_x1_p2_x2_p3")}
            if rank == 0:
                one, _ = slab_leg(local_rank, 0, 1, d, K, Wm, scans=scans)
                r["one_gpu_same_run"] = {"value": one["value"], "ms_per_step": one["ms_per_step"]}
                r["strong_scaling_efficiency"] = r["value"] / one["value"] / world
                r["cut_by_ray_load"]["strong_scaling_efficiency"] = rb["value"] / one["value"] / world
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
    """The oracle (CPU restatement of the reference's path, pinned to the reference's own compiled functions) on the host, single thread,
    first scans of the same sequence; per-stage times under the reference's ScopeTimer checkpoint names, and what its three actors
    (scan thread, raycast thread, background-cluster thread, one core each) would sustain if they overlapped perfectly."""
    from oracle import oracle  # the ONLY use of oracle/ in this file besides --impl reference: the reported CPU baseline
    from vofod_b200 import abi, synth
    p = make_params()
    dirs = synth.sim_lut(W, H)
    o = oracle.Oracle(track_counts=False)
    o.reset(p, VOXEL)
    o.set_sensor(W, H, dirs)
    t_total, n = 0.0, 0
    stage = np.zeros(abi.N_STAGES)
    for k in range(n_scans):
        scan, pose, rp, _ = synth.generate(synth.SCENE_CITY, k, W, H, dirs)
        s = abi.schedule_s1(rp)
        t0 = time.perf_counter()
        o.process_scan(scan, pose, p, s)
        dt = time.perf_counter() - t0
        if k >= timed_from:
            t_total += dt
            n += 1
            stage += o.stage_times()
    o.close()
    stage /= max(n, 1)
    names = ["range", "filtering", "clusterization", "close X far", "vmap update", "raycasting", "raycast vmap update", "classification", "detections",
             "sep bg clusters"]
    table = {nm: round(float(stage[i]), 3) for i, nm in enumerate(names)}
    scan_actor = float(stage[0] + stage[1] + stage[2] + stage[3] + stage[4] + stage[7] + stage[8])   # processMsg (:882-1096)
    ray_actor = float(stage[5] + stage[6])                                                             # raycast_cloud (:1397-1606)
    bg_actor = float(stage[9])                                                                         # updateSeparatedBGClusters (:1126-1278)
    slowest = max(scan_actor, ray_actor, bg_actor)
    port_value = n / t_total if t_total > 0 else None
    ref_value = reference_nodelet_rate(n_scans, timed_from)
    return {"value": ref_value if ref_value else port_value, "unit": "scans/s", "cores": 1, "kind": "reference" if ref_value else "port",
            "sample": f"scans {timed_from}..{n_scans - 1} of the same sequence, schedule S1, g++ -O3 -DNDEBUG (reference flags), 1 thread; host has {os.cpu_count()} cpus"
                      + ("; kind reference = the reference's own per-scan functions of vofod_nodelet.cpp + voxel_map.cpp + voxel_grid_*.cpp compiled from its sources "
                         "(oracle/_ref; PCL / Eigen calls through the stand-ins of oracle/shim)" if ref_value else ""),
            "port_value": port_value,
            "stage_ms_per_scan": table,
            "three_actors_on_three_cores": {"value": 1e3 / slowest if slowest > 0 else None, "unit": "scans/s", "cores": 3,
                                            "actor_ms": {"scan thread": round(scan_actor, 3), "raycast thread": round(ray_actor, 3), "bg-cluster thread": round(bg_actor, 3)},
                                            "note": "upper bound derived from the per-stage times above: the reference's three actors on one core each, perfectly "
                                                    "overlapped, are paced by the slowest one (the raycast thread also skips scans while it is busy, :952-957)"}}


def reference_nodelet_rate(n_scans, timed_from, first=0):
    """scans/s of oracle/_ref = the reference's own compiled functions (None when the library is not in this checkout)"""
    try:
        from oracle import ref
        if not ref.available():
            return None
        from vofod_b200 import abi, synth
        p = make_params()
        dirs = synth.sim_lut(W, H)
        rn = ref.RefNodelet()
        rn.reset(p, VOXEL)
        rn.set_sensor(W, H)
        t_total, n = 0.0, 0
        for k in range(first, first + n_scans):
            scan, pose, rp, _ = synth.generate(synth.SCENE_CITY, k, W, H, dirs)
            s = abi.schedule_s1(rp)
            t0 = time.perf_counter()
            rn.process_scan(scan, pose, p, s)
            dt = time.perf_counter() - t0
            if k - first >= timed_from:
                t_total += dt
                n += 1
        rn.close()
        return n / t_total if t_total > 0 else None
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from vofod_b200 import abi, synth
    K, Wm = args.steps, args.warmup
    if K + Wm > 60:  # bounded sample: ~0.65 s (port) + ~0.8 s (the reference's own code) of CPU work per scan
        K, Wm = min(K, 50), min(Wm, 10)
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
    port_val = K / t_total
    # the reference's OWN compiled code (oracle/_ref) over the same scans, when this checkout has it
    ref_val = reference_nodelet_rate(K + Wm, Wm)
    val = ref_val if ref_val else port_val
    kind = "reference" if ref_val else "port"
    sample = (f"{K} scans (after {Wm} warm-up scans) of the same sequence, schedule S1 fully serial on 1 thread — the reference runs each of its actors "
              f"(scan, raycast, background clusters) on a single thread (pointcloud_threads: 1); host has {os.cpu_count()} cpus"
              + ("; the reference's own per-scan functions compiled from its sources (oracle/_ref), PCL / Eigen calls through oracle/shim" if ref_val else ""))
    out = {"impl": "reference", "metric": "scans/s", "value": val, "unit": "scans/s", "n_gpus": args.gpus, "steps": K, "warmup": Wm,
           "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": "the reference's own vofod_nodelet.cpp member functions / voxel_map.cpp / voxel_grid_*.cpp compiled from its sources "
                                                     "(oracle/_ref) when present, else the CPU oracle port; the ROS nodelet itself needs ROS/PCL/Eigen and cannot be built here"},
           "gvoxel_traversals_per_s": trav / t_ray / 1e9 if t_ray > 0 else None, "port_value": port_val,
           "cpu_baseline": {"value": val, "unit": "scans/s", "cores": 1, "kind": kind, "sample": sample},
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
    ap.add_argument("--no-defer-sep", action="store_true", help="run the separated-background pass at the end of its own step (round-1 behaviour)")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg3 (scene with UAVs / detections) sub-record")
    ap.add_argument("--no-numa-pin", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--slab-steps", type=int, default=20)
    ap.add_argument("--slab-warmup", type=int, default=30, help="past the bootstrap of the background (far clusters of 10^4 points until then)")
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
