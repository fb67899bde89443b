"""k_raycast_accumulate / k_raycast_apply on the large map (cfg5: 0.25 m voxels, 500 x 500 x 100 m) on ONE GPU, unsharded, per raycast.max_distance:
kernel-by-kernel stage events, the shipped loop (exp 0) against round 1's (exp 3).
    python tools/exp_raycast_long.py [dmax ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vofod_b200 import abi, capi, synth  # noqa: E402

W, H = 2048, 128
d = synth.sim_lut(W, H)
out = {}
for dmax in [float(x) for x in sys.argv[1:]] or [20.0, 200.0]:
    p = abi.default_params()
    for i, (o, s) in enumerate(zip((0., 0., -1.25), (500., 500., 100.))):
        p.oparea_offset[i] = o
        p.oparea_size[i] = s
    p.raycast_max_distance = dmax
    v = capi.Vofod(0)
    v.set_option(abi.OPT_GRAPH, 0)
    v.reset(p, 0.25)
    v.set_sensor(W, H, d)
    scans = [synth.generate(synth.SCENE_CITY, k, W, H, d, 2.5) for k in range(30, 33)]
    for name, exp, spread in (("shipped", 0, 64), ("round1_loop", 3, 64), ("spread32", 0, 32), ("spread48", 0, 48), ("spread96", 0, 96), ("shipped_again", 0, 64)):
        v.set_option(abi.OPT_RAYCAST_EXP, exp)
        v.set_option(abi.OPT_RAYCAST_SPREAD, spread)
        ts, ta = [], []
        for (scan, pose, rp, _) in scans * 2:
            s = abi.schedule_s1(rp, do_classify=False, do_sepclusters=False)
            res, _ = v.process_scan(scan, pose, p, s)
            st = v.stage_times()
            ts.append(st['raycasting'])
            ta.append(st.get('raycast vmap update', 0.0))
        out[f"{int(dmax)}m_{name}"] = {"accumulate_ms": round(float(np.median(ts[2:])), 4), "apply_ms": round(float(np.median(ta[2:])), 4), "traversals": int(res.n_traversals)}
    v.close()
print(json.dumps(out))
