"""2+ ranks, one GPU each (torchrun): slab mode through the library's own NCCL path (vofod_comm_init + vofod_slab_process_scan: scan
broadcast, all-reduces and all-gather on the library's stream) against the monolithic CPU oracle — whole schedule S1 including
classification, detections and the separated-background pass.  Every rank checks its result record, voxels, labels, detections and ITS
slab's storage box (own range + halo) of the score grid bit for bit after every scan; rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/check_slab_nccl.py

tests/test_parity_gpu.py::test_slab_mode_over_nccl runs it when the box has at least 2 GPUs.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from harness import Sensor, small_params  # noqa: E402
from oracle import oracle  # noqa: E402  (checker)
from vofod_b200 import abi, capi, slab  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    v = capi.Vofod(local)
    w = slab.SlabWorker(v, p, vs, (sensor.W, sensor.H), sensor.dirs, rank, world, halo=6)
    cpu = oracle.Oracle()
    cpu.reset(p, vs)
    cpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
    sx, sy, sz = list(cpu.map_info().sizes)
    ok, n_scans, n_det, why = True, 30, 0, ""
    for k in range(n_scans):
        scan, pose, rp, _ = sensor.scan(1, k)           # every rank can generate the scan; only rank 0 FEEDS it
        s = abi.schedule_s1(rp)
        res, dg = w.step(scan, pose, s)
        cpu.set_modes(True, True, v.raycast_frac_bits() or 24)
        want, dc = cpu.process_scan(scan, pose, p, s)
        n_det += len(dc)
        mi = v.map_info()
        x0, nx = mi.storage_lo[0], mi.storage_size[0]
        full = cpu.map_download().reshape(sz, sy, sx)
        vg, lg, ig = v.last_voxels()
        vc, lc, ic = cpu.last_voxels()
        checks = {"result": res.as_dict() == want.as_dict(), "map": np.array_equal(v.map_download().reshape(sz, sy, nx), full[:, :, x0:x0 + nx], equal_nan=True),
                  "voxels": vg.tobytes() == vc.tobytes(), "labels": np.array_equal(lg, lc) and np.array_equal(ig, ic),
                  "detections": len(dg) == len(dc) and np.array_equal(dg["id"], dc["id"]) and np.array_equal(dg["label"], dc["label"])
                  and np.allclose(dg["position"], dc["position"], rtol=1e-5, atol=1e-6) and np.allclose(dg["confidence"], dc["confidence"], rtol=1e-5)}
        if not all(checks.values()) and ok:
            why = f"rank {rank} scan {k}: " + ", ".join(n for n, c in checks.items() if not c) + f" {res.as_dict()} vs {want.as_dict()}"
        ok = ok and all(checks.values())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if why:
        print(why, file=sys.stderr, flush=True)
    if rank == 0:
        print(json.dumps({"check": "slab mode over NCCL (library path) vs monolithic oracle: schedule S1 incl. classification, detections, sepclusters", "world": world,
                          "scans": n_scans, "halo": 6, "detections": n_det, "all_slabs_bit_exact": bool(flag.item())}), flush=True)
    dist.barrier()
    v.close()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
