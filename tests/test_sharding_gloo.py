"""CPU test of the multi-GPU host logic with world_size 2 over gloo: slices tile the batch, the final gather
reassembles per-rank results in scenario order, timing reductions take the max over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mujoco_lip_mpc_simulation_b200.sharding import gather_to_rank0, max_over_ranks, shard_bounds, sum_over_ranks


def test_shard_bounds_tile_the_batch():
    for B in (0, 1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(B, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(B, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0])   # "result" rows
    status = torch.arange(lo, hi, dtype=torch.int32) % 3
    full = gather_to_rank0(local, B)
    st = gather_to_rank0(status, B)
    t = max_over_ranks(1.0 + rank, "cpu")
    s = sum_over_ranks(float(hi - lo), "cpu")
    assert t == float(world) and s == float(B)
    if rank == 0:
        assert torch.equal(full, torch.arange(B, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0]))
        assert torch.equal(st, torch.arange(B, dtype=torch.int32) % 3)
    else:
        assert full is None and st is None
    dist.destroy_process_group()


def test_gather_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 1001), nprocs=2, join=True)
