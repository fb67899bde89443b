"""not gpu: the N>1 host logic (stream assignment, slab partition, max-over-ranks timing) with world_size 2 on gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vofod_b200 import multi


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank r "measures" (r+1)*10 ms for 80 scans
        tmax, usum = multi.aggregate((rank + 1) * 10.0, 80.0)
        lo, hi = multi.partition(401, rank, world)
        idx = [multi.stream_scan_index(rank, k) for k in (0, 19, 20, 99)]
        q.put((rank, tmax, usum, lo, hi, idx))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_aggregation_and_partition():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tmax, usum, lo, hi, idx in out:
        assert tmax == 20.0 and usum == 160.0           # max over ranks, sum over ranks
    assert (out[0][3], out[0][4]) == (0, 201) and (out[1][3], out[1][4]) == (201, 401)  # contiguous cover of 401 cells
    assert out[0][5] == [0, 19, 20, 99] and out[1][5] == [0, 19, 1020, 1099]           # shared take-off, own trajectory


def test_partition_covers_everything():
    for n in (1, 7, 401, 2001):
        for world in (1, 2, 4, 8):
            edges = [multi.partition(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_partition_by_ray_load():
    """the load-balanced slab cut: contiguous cover, narrow slabs around the sensor, and a far more even share of the ray work"""
    import numpy as np
    n, world, centre, reach = 2001, 8, 1080.0, 800.0
    cuts = multi.partition_by_ray_load(n, world, centre, reach)
    assert cuts[0][0] == 0 and cuts[-1][1] == n and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
    assert all(hi - lo >= 8 for lo, hi in cuts)
    x = np.arange(n) + 0.5
    dens = np.where(np.abs(x - centre) < reach, np.log(reach / np.maximum(np.abs(x - centre), 0.5)), 0.0)
    share = lambda c: np.array([dens[lo:hi].sum() for lo, hi in c]) / dens.sum()
    uniform = [multi.partition(n, r, world) for r in range(world)]
    assert share(cuts).max() < 0.2 < 0.3 < share(uniform).max()
    assert multi.partition_by_ray_load(401, 1, 200.0, 80.0) == [(0, 401)]
