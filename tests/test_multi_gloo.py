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


def _dry_run(world, extra, timeout=240):
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(root, "tests", "dry", "bench_dry_run.py"), "--gpus", str(world)] + extra
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        pytest.fail(f"bench.py --gpus {world} {' '.join(extra)} did not finish: the ranks' collective sequences do not pair up (dry run, gloo)")
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, (r.returncode, r.stdout[-1000:], r.stderr[-2000:])
    return json.loads(lines[0])


@pytest.mark.parametrize("world", [2, 3])
def test_bench_control_flow_under_torchrun(world):
    """`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N` as the driver launches it, with the CUDA library / runtime and
    the scan generator replaced by stand-ins (tests/dry/bench_dry_run.py) and gloo for NCCL: every rank issues the same collectives in the
    same order (the run ends instead of hanging), rank 0 prints exactly one JSON line with the contract's keys and the sub-records, all
    ranks exit 0.  Round 2's first 8-GPU run of the final bench hung on a rank-0-only dist.barrier; this is the guard."""
    d = _dry_run(world, ["--steps", "3", "--warmup", "3", "--slab-steps", "2", "--slab-warmup", "2"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "gpu_launches", "roofline", "clocks", "slab_cfg5", "cfg3"):
        assert key in d, key
    assert d["n_gpus"] == world and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    slab = d["slab_cfg5"]
    assert "error" not in slab and slab["20m"]["n_slabs"] == world and slab["200m"]["n_slabs"] == world
    assert "one_gpu_same_run" in slab["200m"] and "cut_by_ray_load" in slab["200m"]
    assert len(slab["20m"]["ms_by_rank_bcast_p0_x0_p1_x1_p2_x2_p3"]) == world


def test_bench_slab_mode_control_flow_under_torchrun():
    d = _dry_run(2, ["--mode", "slab", "--raycast-max", "200", "--steps", "2", "--warmup", "2"])
    assert d["mode"] == "slab" and d["n_gpus"] == 2 and d["scaling"] == "strong" and d["slab"]["n_slabs"] == 2


def test_bench_headline_survives_a_stuck_slab_sub_record(monkeypatch):
    """one rank never returns from the cfg5 slab sub-record: after VOFOD_BENCH_SLAB_LIMIT_S every rank abandons it, rank 0 still prints the
    headline line (with the reason in place of the sub-record) and all ranks exit 0"""
    monkeypatch.setenv("DRY_RUN_HANG_IN_SLAB", "1")
    monkeypatch.setenv("VOFOD_BENCH_SLAB_LIMIT_S", "12")
    d = _dry_run(2, ["--steps", "3", "--warmup", "3", "--slab-steps", "2", "--slab-warmup", "2"], timeout=120)
    assert d["n_gpus"] == 2 and d["value"] > 0 and "error" in d["slab_cfg5"]


def test_reference_arm_under_torchrun():
    """`bench.py --impl reference --gpus N` launched like the GPU arm: rank 0 alone runs the reference's CPU path and prints the line, the
    other ranks leave at once with status 0 (the real thing, no stand-ins: the arm needs no GPU)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scans/s" and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
