"""GPU test of config 5's multi-GPU path at world size 2 (skipped on a one-GPU box): every rank draws the scenario set on its
GPU, rolls out its batch slice (dcbf_rollout) and the result is gathered over NCCL -- the gathered tensors must equal, bit for
bit, what one GPU computes for the whole set (MPC_LIP_sig_step.py:565-575 batched)."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

B, STEPS, SEED, NF = 8192, 6, 3, 256


def _single():
    from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
    from mujoco_lip_mpc_simulation_b200.sharding import rollout_shard, shard_inputs
    s = DcbfSolver("sig_step", device=0)
    r = rollout_shard(s, STEPS, shard_inputs(s, B, SEED, NF))
    torch.cuda.synchronize()
    return r


def _worker(rank, world, port, path):
    import torch.distributed as dist
    from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
    from mujoco_lip_mpc_simulation_b200.sharding import gather_rollout, rollout_shard, shard_inputs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s = DcbfSolver("sig_step", device=rank)
    inp = shard_inputs(s, B, SEED, NF, rank=rank, world=world)
    assert inp["hi"] - inp["lo"] == B // world
    g = gather_rollout(rollout_shard(s, STEPS, inp), B)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({k: v.cpu() for k, v in g.items()}, path)
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_rollout_world2_equals_single_gpu(tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("the gpu tests need a CUDA device; there is no CPU fallback")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    path = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, port, path), nprocs=2, join=True)
    g = torch.load(path)
    one = _single()
    assert torch.equal(g["x_final"], one["x_final"].cpu())
    assert torch.equal(g["steps_done"], one["steps_done"].cpu()) and torch.equal(g["n_infeasible"], one["n_infeasible"].cpu())
    assert int(g["steps_done"].sum()) > B      # the scenarios did walk


def test_sharded_rollout_single_rank_path():
    """world size 1: the same helpers without a process group (what bench.py runs at N = 1)"""
    if not torch.cuda.is_available():
        pytest.fail("the gpu tests need a CUDA device; there is no CPU fallback")
    from mujoco_lip_mpc_simulation_b200.sharding import gather_rollout
    one = _single()
    g = gather_rollout(one, B)
    assert torch.equal(g["x_final"], one["x_final"]) and torch.equal(g["steps_done"], one["steps_done"])
    assert int(one["steps_done"].min()) >= 1 and int(one["steps_done"].max()) <= STEPS
