"""2+ ranks, one GPU each (torchrun): slab mode through the NCCL driver (vofod_b200/slab.py: scan broadcast + all-reduce of the
exchange buffers) against the monolithic CPU oracle.  Every rank checks ITS slab's storage box (own range + halo) bit for bit
after every scan; rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/check_slab_nccl.py
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
    w = slab.SlabWorker(v, p, vs, (sensor.W, sensor.H), sensor.dirs, rank, world, halo=8)
    cpu = oracle.Oracle()
    cpu.reset(p, vs)
    cpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
    sx, sy, sz = list(cpu.map_info().sizes)
    N = sensor.W * sensor.H
    pinned = torch.empty(N * abi.PT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    ok, n_scans = True, 24
    for k in range(n_scans):
        scan, pose, rp, _ = sensor.scan(0, k)           # every rank can generate the scan; only rank 0 FEEDS it
        if rank == 0:
            pinned.numpy().view(abi.PT_DTYPE)[:] = scan
        res = w.step(pinned if rank == 0 else None, pose if rank == 0 else None, rp if rank == 0 else None)
        s = abi.schedule_s1(rp, do_classify=False, do_sepclusters=False)
        cpu.set_modes(True, True, v.raycast_frac_bits() or 24)
        want, _ = cpu.process_scan(scan, pose, p, s)
        mi = v.map_info()
        x0, nx = mi.storage_lo[0], mi.storage_size[0]
        full = cpu.map_download().reshape(sz, sy, sx)
        ok = ok and res.as_dict() == want.as_dict() and np.array_equal(v.map_download().reshape(sz, sy, nx), full[:, :, x0:x0 + nx])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "slab mode over NCCL vs monolithic oracle", "world": world, "scans": n_scans, "halo": 8,
                          "all_slabs_bit_exact": bool(flag.item())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
