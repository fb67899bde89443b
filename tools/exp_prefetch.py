import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vofod_b200 import abi, capi, synth
W, H = 2048, 128
d = synth.sim_lut(W, H)
p = abi.default_params()
for i, (o, s) in enumerate(zip((0., 0., -1.25), (200., 200., 80.))):
    p.oparea_offset[i] = o; p.oparea_size[i] = s
v = capi.Vofod(0); v.set_sensor(W, H, d)
n = 60
pinned = torch.empty((n, W * H * 20), dtype=torch.uint8, pin_memory=True)
hs = pinned.numpy().view(abi.PT_DTYPE).reshape(n, W * H)
meta = []
for k in range(n):
    _, pose, rp, _ = synth.generate(0, k, W, H, d, 1.0, out=hs[k]); meta.append((pose, abi.schedule_s1(rp)))
for mode in ("plain", "prefetch", "plain", "prefetch"):
    v.reset(p, 0.5)
    torch.cuda.synchronize()
    ts = []
    for k in range(n):
        t0 = time.perf_counter()
        if mode == "prefetch" and k + 1 < n:
            v.prefetch_scan(hs[k + 1])
        v.process_scan(hs[k], meta[k][0], p, meta[k][1])
        ts.append(time.perf_counter() - t0)
    print(mode, "median wall ms/scan", round(1e3 * float(np.median(ts[30:])), 4), v.stats(), flush=True)
