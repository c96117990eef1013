"""Scenario sharding across the GPUs of one box.

Every scenario is an independent NLP / closed-loop chain, so the path shards with NO data-path collective:
rank g owns the contiguous slice [g*B/G, (g+1)*B/G).  The only communication is the final gather of small result
tensors (status / iteration counts / plans) to rank 0, done with torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of rank `rank`; sizes differ by at most one; slices tile [0, B) exactly."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_to_rank0(t: torch.Tensor, B: int) -> torch.Tensor | None:
    """Gather per-rank result slices (first dim = scenarios of shard_bounds) into one [B, ...] tensor on rank 0."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(B, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != 0:
        return None
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
