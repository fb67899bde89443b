"""One slab of an N-slab run alone on one GPU (no NCCL: the exchange buffers are left as this slab computed them, the gathered background
list holds this slab's part only), so that its kernels can be listed with ncu — profiling aid for bench.py --mode slab.
    python tools/exp_slab_rank.py RANK WORLD RAYCAST_MAX [N_SCANS] [balanced]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vofod_b200 import abi, capi, multi, synth  # noqa: E402

rank, world, dmax = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
n_scans = int(sys.argv[4]) if len(sys.argv) > 4 else 40
balanced = len(sys.argv) > 5
W, H = 2048, 128
p = abi.default_params()
for i, (o, sz) in enumerate(zip((0.0, 0.0, -1.25), (500.0, 500.0, 100.0))):
    p.oparea_offset[i] = o
    p.oparea_size[i] = sz
p.raycast_max_distance = dmax
dirs = synth.sim_lut(W, H)
v = capi.Vofod(0)
v.reset(p, 0.25)
cuts = multi.partition_by_ray_load(2001, world, (12.0 + 250.0) / 0.25, dmax / 0.25 + 120.0) if balanced else None
lo, hi = cuts[rank] if cuts else multi.partition(2001, rank, world)
v.set_slab(0, lo, hi, 16)
v.map_set_to(abi.MAP_SCORE, p.score_init)
v.set_sensor(W, H, dirs)
v.slab_set_world(rank, world)
t_phase = np.zeros(4)
for k in range(n_scans):
    scan, pose, rp, _ = synth.generate(synth.SCENE_CITY, k, W, H, dirs, 2.5)
    s = abi.schedule_s1(rp)
    for ph in range(4):
        t0 = time.perf_counter()
        out = v.slab_phase(ph, scan, pose, p, s) if ph == 0 else v.slab_phase(ph)
        if ph < 3:
            v.synchronize()
        if ph == 2:
            for x in v.slab_exchanges(2):  # this slab's own list into its place of the gathered buffer
                own = v.dev_read(x.buf, np.uint32, x.count)
                v.dev_write(x.gather_out + rank * x.count * 4, own)
        if k >= n_scans - 10:
            t_phase[ph] += time.perf_counter() - t0
print(json.dumps({"rank": rank, "world": world, "own": [lo, hi], "raycast_max": dmax, "host_ms_per_phase_last10": (t_phase * 100).round(3).tolist(),
                  "result": out[1].as_dict()}))
