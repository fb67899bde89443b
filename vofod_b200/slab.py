"""Large-map mode (BASELINE.json configs[4]): the grid is cut into spatial slabs along x, one slab per GPU / process
(vofod_b200/csrc/slab.cu says what is sharded, what is replicated and what crosses slabs).

Two drivers over the same four-phase scan of the library:
  * SlabWorker     — one process per GPU: the library's own NCCL path (vofod_comm_init + vofod_slab_process_scan): scan broadcast,
                     all-reduces and all-gather on the library's stream, one host wait per scan.  torch.distributed only carries the
                     128-byte NCCL id to the other ranks.
  * run_emulated   — several slabs as several contexts on ONE device, the exchange buffers combined by a host loop (tests)."""
import ctypes as C

import numpy as np

from . import abi, multi


def make_slabs(ctxs, params, voxel_size, sensor_wh, dirs, halo, axis=0, offs=None, mask=None):
    """reset every context, give it its slab of the grid and the sensor"""
    W, H = sensor_wh
    n = len(ctxs)
    for r, g in enumerate(ctxs):
        g.reset(params, voxel_size)
        sizes = list(g.map_info().sizes)
        lo, hi = multi.partition(sizes[axis], r, n)
        g.set_slab(axis, lo, hi, halo)
        g.map_set_to(abi.MAP_SCORE, params.score_init)
        g.set_sensor(W, H, dirs, offs, mask)
        g.slab_set_world(r, n)


_NP = {abi.XCHG_SUM_U64: np.uint64, abi.XCHG_MAX_I32: np.int32, abi.XCHG_SUM_U32: np.uint32, abi.XCHG_GATHER_U32: np.uint32}


def _combine(ctxs, phase):
    lists = [g.slab_exchanges(phase) for g in ctxs]
    for j in range(len(lists[0])):
        kind, count = lists[0][j].kind, lists[0][j].count
        assert all(x[j].kind == kind and x[j].count == count for x in lists)
        parts = [g.dev_read(x[j].buf, _NP[kind], count) for g, x in zip(ctxs, lists)]
        if kind == abi.XCHG_GATHER_U32:
            allp = np.concatenate(parts)
            for g, x in zip(ctxs, lists):
                g.dev_write(x[j].gather_out, allp)
            continue
        if kind == abi.XCHG_MAX_I32:
            comb = np.maximum.reduce(parts)
        else:
            comb = np.add.reduce(np.stack(parts), axis=0, dtype=_NP[kind])
        for g, x in zip(ctxs, lists):
            g.dev_write(x[j].buf, comb)


def run_emulated(ctxs, scan, pose, params, sched, det_cap=256):
    """one slab-mode scan on every context of `ctxs` (all slabs of one map); returns [(ScanResult, detections)] per slab"""
    for g in ctxs:
        g.slab_phase(0, scan, pose, params, sched)
    _combine(ctxs, 0)
    for g in ctxs:
        g.slab_phase(1)
    _combine(ctxs, 1)
    while True:
        for g in ctxs:
            g.slab_phase(2)
        _combine(ctxs, 2)
        out = [g.slab_phase(3, det_cap=det_cap) for g in ctxs]
        codes = {o[0] for o in out}
        assert len(codes) == 1, codes  # every slab sees the same gathered counts
        if codes == {abi.VOFOD_W_REDO}:
            continue
        return [(o[1], o[2]) for o in out]


class SlabWorker:
    """one slab of the map in this process; rank / world come from torch.distributed (NCCL or gloo: it only ships the NCCL id)"""

    def __init__(self, ctx, params, voxel_size, sensor_wh, dirs, rank, world, halo=16, axis=0, cuts=None):
        """cuts: None = equal-width slabs, or the [lo, hi) list of multi.partition_by_ray_load"""
        self.v, self.p, self.rank, self.world = ctx, params, rank, world
        W, H = sensor_wh
        ctx.reset(params, voxel_size)
        sizes = list(ctx.map_info().sizes)
        lo, hi = cuts[rank] if cuts is not None else multi.partition(sizes[axis], rank, world)
        self.lo, self.hi = lo, hi
        ctx.set_slab(axis, lo, hi, halo)
        ctx.map_set_to(abi.MAP_SCORE, params.score_init)
        ctx.set_sensor(W, H, dirs)
        uid = None
        if world > 1:
            import torch
            import torch.distributed as dist
            buf = np.zeros(128, dtype=np.uint8)
            if rank == 0:
                rc = ctx.lib.vofod_comm_unique_id(buf.ctypes.data_as(C.c_void_p))
                assert rc == 0, "libnccl.so.2 not loadable"
            t = torch.from_numpy(buf)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            uid = t.cpu().numpy().tobytes()
        ctx.comm_init(rank, world, uid)

    def step(self, scan_host, pose, sched, next_scan_host=None):
        """scan_host is read on rank 0 only; pose and schedule are given on every rank.  next_scan_host (rank 0): the scan of the NEXT step,
        announced with vofod_prefetch_scan so that its host->device copy runs beside this step's kernels.  -> (ScanResult, detections)"""
        if self.rank == 0 and next_scan_host is not None:
            self.v.prefetch_scan(next_scan_host)
        return self.v.slab_process_scan(scan_host if self.rank == 0 else None, pose, self.p, sched)

    def close(self):
        self.v.close()
