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


def aggregate(local_ms: float, local_units: float, device=None):
    """-> (max over ranks of local_ms, sum over ranks of local_units).  Whole-job throughput = units / max time."""
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
