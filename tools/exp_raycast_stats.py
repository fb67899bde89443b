"""Warp-step statistics of k_raycast_accumulate on the bench workload (cfg2, steady-state scans): lanes in the loop, distinct voxels
per warp-step (= REDs issued), fast/general loop split; plus the kernel's duration per block size (kernel-by-kernel stage events)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vofod_b200 import abi, capi, synth  # noqa: E402

W, H = 2048, 128
d = synth.sim_lut(W, H)
p = abi.default_params()
for i, (o, s) in enumerate(zip((0., 0., -1.25), (200., 200., 80.))):
    p.oparea_offset[i] = o
    p.oparea_size[i] = s
v = capi.Vofod(0)
v.set_option(abi.OPT_GRAPH, 0)
v.reset(p, 0.5)
v.set_sensor(W, H, d)
first = int(sys.argv[1]) if len(sys.argv) > 1 else 40
scans = [synth.generate(0, k, W, H, d) for k in range(first, first + 4)]
out = {}
for rb in (64, 128, 256):
    v.set_option(abi.OPT_RAYCAST_BLOCK, rb)
    v.reset(p, 0.5)
    ts = []
    for (scan, pose, rp, _) in scans * 3:
        res, _ = v.process_scan(scan, pose, p, abi.schedule_s1(rp))
        ts.append(v.stage_times()['raycasting'])
    out[f"raycasting_ms_rb{rb}"] = [round(float(x), 4) for x in ts[4:]]
    out["traversals"] = int(res.n_traversals)
v.set_option(abi.OPT_RAYCAST_BLOCK, 64)
v.set_option(abi.OPT_RAYCAST_STATS, 1)
v.reset(p, 0.5)
v.raycast_stats()
trav = 0
for (scan, pose, rp, _) in scans:
    res, _ = v.process_scan(scan, pose, p, abi.schedule_s1(rp))
    trav += res.n_traversals
lanes, groups, fast, general, skipped = v.raycast_stats()
ws = int(lanes.sum())
out.update({"scans": len(scans), "traversals_total": int(trav), "warp_steps": ws, "fast_warp_steps": fast, "general_warp_steps": general,
            "mean_lanes_per_warp_step": float((lanes * np.arange(33)).sum() / max(ws, 1)),
            "mean_groups_per_warp_step": float((groups * np.arange(33)).sum() / max(ws, 1)),
            "lanes_hist": lanes.tolist(), "groups_hist": groups.tolist()})
print(json.dumps(out))
