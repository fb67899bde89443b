"""bench.py's control flow on CPU: the CUDA library, the CUDA runtime and the scan generator are replaced by stand-ins that do nothing, the process
group is gloo.  What is left is exactly what this checks — which rank calls which collective in which order (a rank-0-only sub-record that called
dist.barrier made `bench.py --gpus N` hang for every N > 1 in round 2), that rank 0 prints ONE JSON line with the contract's keys, and that
every rank leaves with status 0.  Launched by tests/test_multi_gloo.py under torch.distributed.run; NOT a measurement of anything."""
import contextlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vofod_b200 import abi, capi, synth  # noqa: E402


# ---- torch.cuda / device tensors -------------------------------------------------------------------------------------------------------
class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 0.25


class _Stream:
    def __init__(self, *a, **k):
        pass


def _strip(kw):
    kw.pop("pin_memory", None)
    if str(kw.get("device", "")).startswith("cuda"):
        kw.pop("device")
    return kw


for name in ("empty", "tensor", "zeros"):
    orig = getattr(torch, name)
    setattr(torch, name, (lambda o: lambda *a, **k: o(*a, **_strip(k)))(orig))
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.empty_cache = lambda *a, **k: None
torch.cuda.Event = _Event
torch.cuda.ExternalStream = _Stream
torch.cuda.stream = lambda s: contextlib.nullcontext()
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init("gloo")


# ---- the library -----------------------------------------------------------------------------------------------------------------------
class _Lib:
    def vofod_comm_unique_id(self, buf):
        return 0


class _MapInfo:
    sizes = [2001, 2001, 401]
    storage_size = [2001, 2001, 401]


class FakeVofod:
    def __init__(self, device):
        self.lib = _Lib()
        self.n = 0

    def _res(self):
        r = abi.ScanResult()
        r.n_traversals = 1000
        r.n_voxels = 10
        return r

    def __getattr__(self, name):  # everything that only has side effects on the device
        return lambda *a, **k: None

    def stream(self):
        return 0

    def kernel_launches(self):
        self.n += 32
        return self.n

    def stage_times(self):
        return {n: 0.01 for n in ("range", "filtering", "clusterization", "close X far", "vmap update", "raycasting", "raycast vmap update", "classification",
                                  "detections", "sep bg clusters", "readback", "total")}

    def stats(self):
        return {"graph_replays": 0, "captures": 0, "failed_captures": 0, "eager_scans": 0, "last_capture_error": 0, "prefetch_hits": 0}

    def map_info(self):
        return _MapInfo()

    def slab_times(self):
        return np.full(8, 0.1, dtype=np.float32)

    def process_scan(self, scan, pose, p, s, **k):
        return self._res(), np.zeros(0, dtype=abi.DETECTION_DTYPE)

    def process_scan_resident(self, slot, pose, p, s, **k):
        return self._res(), np.zeros(0, dtype=abi.DETECTION_DTYPE)

    def slab_process_scan(self, scan, pose, p, s, **k):
        if os.environ.get("DRY_RUN_HANG_IN_SLAB") and int(os.environ.get("RANK", "0")) == 1:
            import time
            time.sleep(3600)  # a stuck rank: the bench's deadline guard must still get the headline line out
        return self._res(), np.zeros(0, dtype=abi.DETECTION_DTYPE)

    def process_scan_batch(self, scans, poses, p, scheds, **k):
        return [self._res() for _ in scans], len(scans)


capi.Vofod = FakeVofod


def _generate(scene_id, k, W, H, dirs, map_scale=1.0, out=None):
    if out is None:
        out = np.zeros(W * H, dtype=abi.PT_DTYPE)
    return out, abi.Pose(), np.zeros(3, dtype=np.float32), np.zeros((3, 3), dtype=np.float32)


synth.generate = _generate
synth.sim_lut = lambda W, H: np.zeros((3, W * H), dtype=np.float32)

import bench  # noqa: E402

bench.pin_to_gpu_numa_node = lambda r: {"applied": False, "dry_run": True}
bench.cpu_baseline = lambda *a, **k: {"dry_run": True}

if __name__ == "__main__":
    bench.main()
