"""Where the time of k_raycast_accumulate goes on the bench workload (cfg2, steady-state scans): the kernel as it ships, the same without
the RED instruction (DDA + match / redux), and the DDA alone — kernel-by-kernel stage events, ms per launch."""
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
scans = [synth.generate(0, k, W, H, d) for k in range(40, 44)]
out = {}
for name, exp in (("full", 0), ("round1_loop_56regs", 3), ("no_red_round1_loop", 1), ("dda_only_round1_loop", 2), ("full_again", 0)):
    v.set_option(abi.OPT_RAYCAST_EXP, exp)
    ts = []
    for (scan, pose, rp, _) in scans * 3:
        s = abi.schedule_s1(rp, do_classify=False, do_sepclusters=False)
        res, _ = v.process_scan(scan, pose, p, s)
        ts.append(v.stage_times()['raycasting'])
    out[name + "_ms"] = round(float(np.median(ts[4:])), 4)
    out["traversals"] = int(res.n_traversals) if exp not in (1, 2) else out.get("traversals")
print(json.dumps(out))
