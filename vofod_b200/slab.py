"""Large-map mode (BASELINE.json configs[4]): the grid is cut into spatial slabs along x, one slab per GPU / process.

Per scan: rank 0 holds the scan -> NCCL broadcast of the packed scan (5.2 MB) and of the pose / seed record -> every rank runs
vofod_slab_scan_begin on the broadcast buffer (device pointer) -> NCCL all-reduce of the two exchange buffers (SUM of the
8-byte background count, MAX of the per-cluster close flags), enqueued on the library's own stream so that nothing waits on the
host -> vofod_slab_scan_end.  See vofod_b200/csrc/slab.cu for why halos need no exchange."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import abi, multi


class _Raw:
    """__cuda_array_interface__ view of a raw device pointer, so that torch can wrap library-owned memory."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class SlabWorker:
    def __init__(self, ctx, params, voxel_size, sensor_wh, dirs, rank, world, halo=16, axis=0):
        self.v, self.p, self.rank, self.world = ctx, params, rank, world
        W, H = sensor_wh
        ctx.reset(params, voxel_size)
        sizes = list(ctx.map_info().sizes)
        lo, hi = multi.partition(sizes[axis], rank, world)
        ctx.set_slab(axis, lo, hi, halo)
        ctx.map_set_to(abi.MAP_SCORE, params.score_init)
        ctx.set_sensor(W, H, dirs)
        self.n = W * H
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.dev)
        self.scan_dev = torch.empty(self.n * abi.PT_DTYPE.itemsize, dtype=torch.uint8, device=self.dev)
        self.meta = torch.empty(16, dtype=torch.float32, device=self.dev)  # R[9], t[3], range_pt[3], pad

    def step(self, scan_host_pinned, pose, range_pt, do_raycast=True):
        """scan_host_pinned / pose / range_pt are read on rank 0 only.  Returns the vofod_scan_result of this slab."""
        with torch.cuda.stream(self.stream):
            if self.rank == 0:
                self.scan_dev.copy_(scan_host_pinned, non_blocking=True)
                m = np.zeros(16, dtype=np.float32)
                m[:9], m[9:12], m[12:15] = list(pose.R), list(pose.t), list(range_pt)
                self.meta.copy_(torch.from_numpy(m))
            if self.world > 1:
                dist.broadcast(self.scan_dev, src=0)
                dist.broadcast(self.meta, src=0)
            m = self.meta.cpu().numpy()
            pose = abi.Pose.from_arrays(m[:9], m[9:12])
            s = abi.schedule_s1(m[12:15], do_raycast=do_raycast, do_classify=False, do_sepclusters=False)
            self.v.slab_scan_begin(None, pose, self.p, s, device_ptr=self.scan_dev.data_ptr())
            if self.world > 1:
                p_nbg, p_close, n = self.v.slab_exchange_buffers()
                nbg = torch.as_tensor(_Raw(p_nbg, 1, "<i8"), device=self.dev)
                close = torch.as_tensor(_Raw(p_close, n, "<i4"), device=self.dev)
                dist.all_reduce(nbg, op=dist.ReduceOp.SUM)
                dist.all_reduce(close, op=dist.ReduceOp.MAX)
            return self.v.slab_scan_end(self.p, s)

    def close(self):
        """Release every torch object that lives on the library's stream BEFORE the context (and with it the stream) goes away."""
        torch.cuda.synchronize()
        self.scan_dev = None
        self.meta = None
        self.stream = None
        torch.cuda.empty_cache()
        self.v.close()
