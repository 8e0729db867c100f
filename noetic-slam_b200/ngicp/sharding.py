"""Host-side partitioning of independent work units over ranks (one process per GPU).

The hot path has no exchange step: a scan-to-submap alignment runs on one GPU, and the batched
workloads are independent units — keyframes for the bulk covariance build (the reference computes
covariances per scan and concatenates per submap: src/dlio/src/nano_gicp/nano_gicp.cc:174-181,
src/dlio/src/dlio/odom.cc:1719-1729) and whole sequences for multi-sequence registration. So ranks
only share (a) which units they own and (b) the max-over-ranks time and the summed unit count for
reporting. torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries those two scalars; there is
no data-path collective.
"""
from __future__ import annotations

import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process => (0, 0, 1))."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def units_for_rank(n_units: int, rank: int, world_size: int) -> list[int]:
    """Round-robin ownership: unit i -> rank i mod G. Every unit has exactly one owner."""
    return list(range(rank, n_units, world_size))


def keyframe_segments(bounds, owned: list[int]):
    """Offsets of the owned keyframes once they are packed back to back, plus the (start, end) slices
    to gather them from the full cloud."""
    slices = [(int(bounds[i]), int(bounds[i + 1])) for i in owned]
    off = [0]
    for s, e in slices:
        off.append(off[-1] + (e - s))
    return off, slices


def reduce_job(local_seconds: float, local_units: float, device=None):
    """(max over ranks of the time, sum over ranks of the units). Works without a process group."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return local_seconds, local_units
    if not (dist.is_available() and dist.is_initialized()):
        return local_seconds, local_units
    t = torch.tensor([local_seconds], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
