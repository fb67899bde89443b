"""Host-side logic of the multi-GPU modes (one process per GPU, torch.distributed for the plumbing).

Mode 1 — independent scan streams (BASELINE.json configs[3]): rank r owns stream r; nothing crosses ranks on the data
path.  The only collectives are the timing barrier and the max / sum reductions of the bench numbers, which work on any
backend (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def stream_scan_index(rank: int, k: int, takeoff_scans: int = 20, stride: int = 1000) -> int:
    """Scan index of step k for the stream of `rank`: every stream bootstraps with the same take-off (scan indices
    0..takeoff_scans-1), then flies its own part of the trajectory (rank*stride + k)."""
    return k if (rank == 0 or k < takeoff_scans) else rank * stride + k


def partition(n_items: int, rank: int, world: int):
    """Contiguous block partition [lo, hi) of n_items over `world` ranks (used for slab extents and stream lists)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_by_ray_load(n_items: int, world: int, centre: float, reach: float, min_width: int = 8, uniform_share: float = 0.15):
    """Slab boundaries along one axis that even out the RAY work instead of the cell count.  Rays start at one point (the sensor, at cell
    `centre` of this axis) and every ray adds work uniformly along its length, so the traversal density falls off as 1/r^2 around the
    sensor; its marginal along one axis is ~ ln(reach / |x - centre|) for |x - centre| < reach: with equal-width slabs the one holding the
    sensor does several times the average work.  The cut equalises  uniform_share * cells + (1 - uniform_share) * ray_density  (the grid
    passes scale with cells, accumulate + apply with the rays).  Returns `world` [lo, hi) pairs covering [0, n_items)."""
    import numpy as np
    x = np.arange(n_items, dtype=np.float64) + 0.5
    d = np.abs(x - centre)
    dens = np.where(d < reach, np.log(reach / np.maximum(d, 0.5)), 0.0)
    if dens.sum() <= 0:
        return [partition(n_items, r, world) for r in range(world)]
    w = uniform_share / n_items + (1.0 - uniform_share) * dens / dens.sum()
    cum = np.concatenate([[0.0], np.cumsum(w)])
    edges = [0]
    for r in range(1, world):
        e = int(np.searchsorted(cum, r / world))
        e = max(e, edges[-1] + min_width)
        edges.append(min(e, n_items - (world - r) * min_width))
    edges.append(n_items)
    return [(edges[r], edges[r + 1]) for r in range(world)]


def aggregate(local_ms: float, local_units: float, device=None):
    """-> (max over ranks of local_ms, sum over ranks of local_units).  Whole-job throughput = units / max time."""
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
