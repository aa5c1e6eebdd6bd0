"""Multi-GPU plumbing of the batched path: independent trajectories (or noise hypotheses) are
sharded by contiguous ranges, one process per GPU; there is no collective inside the data path.
The only exchange is the final gather of per-trajectory ATE statistics (torch.distributed:
NCCL over NVLink on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of [0, total): rank r owns [r*total//world, (r+1)*total//world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    return rank * total // world, (rank + 1) * total // world


def gather_stats(local: torch.Tensor, total_rows: int | None = None, counts: list[int] | None = None) -> torch.Tensor:
    """All-gather row blocks of possibly different heights (shards differ by at most one row) and
    return them concatenated in rank order on every rank.  No-op without a process group.
    ``counts`` (rows per rank, e.g. from shard_range) saves the exchange of the block heights and every
    host synchronisation: the call is then one collective, asynchronous on the current stream."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        rows = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_rows = [torch.zeros_like(rows) for _ in range(world)]
        dist.all_gather(all_rows, rows)
        counts = [int(r.item()) for r in all_rows]
    width = max(counts)
    if local.shape[0] == width:
        padded = local.contiguous()
    else:
        padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    flat = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, padded)
    if all(c == width for c in counts):
        out = flat
    else:
        out = torch.cat([flat[r * width: r * width + c] for r, c in enumerate(counts)])
    if total_rows is not None and out.shape[0] != total_rows:
        raise RuntimeError(f"gathered {out.shape[0]} rows, expected {total_rows}")
    return out


def max_over_ranks(value: float, device) -> float:
    """Timing reduction: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
