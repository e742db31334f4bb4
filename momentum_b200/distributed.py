"""Multi-GPU plumbing of the batched IK path.

IK instances are independent, so the batch is sharded across ranks (one process per GPU) with no
data-path collective; the only communication is one all-reduce of the aggregate statistics
{iterations, sum of final objectives} (SUM) and of the elapsed device time (MAX). torch.distributed is
used as plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition of `total` instances: the first (total % world) ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def aggregate_solve_stats(iterations: float, error_sum: float, elapsed_ms: float, device=None):
    """All-reduce (SUM, SUM, MAX). Returns (total_iterations, total_error, max_elapsed_ms). Works
    un-initialised (single process) too."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(iterations), float(error_sum), float(elapsed_ms)
    s = torch.tensor([float(iterations), float(error_sum)], dtype=torch.float64, device=device)
    m = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(s[0].item()), float(s[1].item()), float(m[0].item())
