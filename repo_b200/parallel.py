"""Data-parallel plumbing for the row-sharded hot path.

Rows (imagine start states, observe batch columns) are independent, so each rank owns a contiguous
shard and the forward path needs no collective.  What does cross ranks: timing (max over ranks),
means over rows (each rank contributes sum and count so uneven shards stay exact), and — with the
backward pass — flat gradient buckets."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, start+count) of n_rows for `rank`; the first n_rows % world ranks get one
    extra row (50 rows over 8 ranks -> 7,7,6,6,6,6,6,6)."""
    base, rem = divmod(n_rows, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def shard_time_major(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """(T, rows, ...) -> this rank's (T, count, ...) slice, contiguous."""
    start, count = shard_rows(x.shape[1], rank, world)
    return x[:, start:start + count].contiguous()


def global_mean(local_values: torch.Tensor) -> torch.Tensor:
    """Mean over all rows of all ranks of a per-row (or per-(t,row)) tensor: SUM-reduce (sum, count)."""
    acc = torch.stack([local_values.sum().double(), torch.tensor(float(local_values.numel()), dtype=torch.float64,
                                                                  device=local_values.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return (acc[0] / acc[1]).to(local_values.dtype)


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_flat(tensors: Iterable[torch.Tensor]) -> None:
    """Sum-all-reduce a parameter group's tensors as ONE flat bucket (one collective per group)."""
    ts: List[torch.Tensor] = [t for t in tensors if t is not None]
    if not ts or not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    flat = torch.cat([t.reshape(-1) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for t in ts:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
