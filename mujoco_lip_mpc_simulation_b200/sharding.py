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
    out = torch.empty((world * width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad)      # one collective (a single NCCL kernel on GPUs)
    if rank != 0:
        return None
    if all(hi - lo == width for lo, hi in sizes):
        return out
    return torch.cat([out[r * width: r * width + hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


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


# ---- config 5: closed-loop rollout of B scenarios sharded over the ranks (MPC_LIP_sig_step.py:565-575 batched) ----------------
def shard_inputs(solver, B: int, seed: int, n_fields: int = 4096, rank: int = 0, world: int = 1) -> dict:
    """This rank's share of a B-scenario set.  The set is a function of (B, seed, n_fields) only: every rank draws the SAME
    obstacle-field pool and the same B start states on its GPU (dcbf_gen_fields / dcbf_gen_states are keyed by field / scenario
    index; a few ms for a million scenarios, nothing touches the host) and keeps the contiguous slice shard_bounds gives it, so
    a gathered result does not depend on the number of ranks.  Installs the pool as the solver's fields.
    Returns dict(lo, hi, x0, goal, leg, field) of device tensors."""
    from . import scenarios
    sc = scenarios.make_batch_device(solver, B, seed=seed, n_fields=n_fields)
    lo, hi = shard_bounds(B, rank, world)
    sl = slice(lo, hi)
    return dict(lo=lo, hi=hi, x0=sc["x0"][sl].contiguous(), goal=sc["goal"][sl].contiguous(), leg=sc["leg"][sl].contiguous(),
                field=sc["field"][sl].contiguous())


def rollout_shard(solver, steps: int, inp: dict) -> dict:
    """closed-loop rollout of this rank's slice (dcbf_rollout: plan -> apply -> warm-started re-plan, no host round trips);
    enqueues on the current stream.  Returns dict(x_final[n,5], steps_done[n], n_infeasible[n], total_iters[n])."""
    return solver.rollout(steps, inp["x0"], inp["goal"], inp["leg"], field=inp["field"], want_traj=False)


def gather_rollout(res: dict, B: int):
    """The final result gather of a sharded rollout: ONE collective over a packed [n, 7] FP64 buffer (x_final, steps_done,
    n_infeasible) -> rank 0 gets dict(x_final[B,5], steps_done[B], n_infeasible[B]), the other ranks None."""
    packed = torch.cat([res["x_final"], res["steps_done"].to(torch.float64)[:, None], res["n_infeasible"].to(torch.float64)[:, None]], dim=1)
    full = gather_to_rank0(packed, B)
    if full is None:
        return None
    return dict(x_final=full[:, :5].contiguous(), steps_done=full[:, 5].to(torch.int32), n_infeasible=full[:, 6].to(torch.int32))
