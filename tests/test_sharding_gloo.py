"""N > 1 host logic on CPU: two gloo ranks each evaluate the loss path on their batch shard (with the CPU oracle
port standing in for the kernels -- tests may use the oracle); mean-of-means and the re-assembled per-sample
gradients must equal the single-process result on the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(B=4, S=2, H=32, W=64, num_scales=2, seed=5, smooth=True, motion="small")


def _loss_and_grads(inp):
    from oracle import torch_port
    depth = [d.clone().requires_grad_(True) for d in inp.depth_pyr]
    poses = inp.poses.clone().requires_grad_(True)
    total, rigid, smooth = torch_port.loss_path(inp.tgt_pyr, inp.src_pyr, depth, poses, inp.K_ms)
    total.backward()
    return total.detach(), [d.grad for d in depth], poses.grad


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from snde_b200.sharding import mean_over_ranks, shard_inputs
    from snde_b200.synthetic import make_inputs
    full = make_inputs(**CFG)
    mine = shard_inputs(full, rank, world)
    assert mine.B == CFG["B"] // world and mine.src_pyr[0].shape[0] == mine.B * mine.S
    loss, gdepth, gposes = _loss_and_grads(mine)
    loss_mean = mean_over_ranks(loss)
    # what DDP does to parameter gradients: average over ranks.  Per-sample inputs (depth, pose) belong to one rank
    # only, so their global-batch gradient is the local one divided by the number of ranks.
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss_mean=loss_mean.numpy(), gposes=(gposes / world).numpy(),
             **{f"gd{s}": (g / world).numpy() for s, g in enumerate(gdepth)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_the_global_batch(tmp_path):
    from snde_b200.sharding import shard_range, unshard_depth_grads
    from snde_b200.synthetic import make_inputs
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    full = make_inputs(**CFG)
    loss, gdepth, gposes = _loss_and_grads(full)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    for p in parts:                                    # every rank logs the same global loss
        assert abs(float(p["loss_mean"]) - float(loss)) < 1e-6 * abs(float(loss))
    got_poses = torch.cat([torch.from_numpy(p["gposes"]) for p in parts], 0)
    assert torch.allclose(got_poses, gposes, rtol=1e-4, atol=1e-7)
    for s in range(CFG["num_scales"]):
        got = unshard_depth_grads([torch.from_numpy(p[f"gd{s}"]) for p in parts], CFG["B"], CFG["S"])
        assert got.shape == gdepth[s].shape
        assert torch.allclose(got, gdepth[s], rtol=1e-4, atol=1e-9)
    assert shard_range(8, 3, 4) == (6, 8)
    with pytest.raises(ValueError):
        shard_range(6, 0, 4)
