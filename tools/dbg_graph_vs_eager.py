import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from vofod_b200 import abi, capi
from harness import Sensor, small_params
gpu = capi.Vofod(0)
sensor = Sensor(512, 32)
p, vs = small_params()
p.background_sufficient_points_ratio = 0.02
outs = []
for graph in (1, 0):
    gpu.set_option(abi.OPT_GRAPH, graph)
    gpu.reset(p, vs)
    gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
    log = []
    for k in range(40):
        scan, pose, rp, _ = sensor.scan(1, k)
        res, dets = gpu.process_scan(scan, pose, p, abi.schedule_s1(rp))
        log.append((res.as_dict(), dets.copy(), gpu.map_download().copy() if k < 14 else None))
    outs.append(log)
for k in range(40):
    a, b = outs[0][k], outs[1][k]
    if a[0] != b[0]:
        print(k, "res differ", a[0], b[0]); break
    if a[1].tobytes() != b[1].tobytes():
        print(k, "dets differ")
        for f in a[1].dtype.names:
            if not np.array_equal(a[1][f], b[1][f]):
                print("  field", f, a[1][f], b[1][f])
        if a[2] is not None:
            print("  maps equal:", np.array_equal(a[2], b[2]))
            if k: print("  prev maps equal:", np.array_equal(outs[0][k-1][2], outs[1][k-1][2]))
        break
else:
    print("all equal")
