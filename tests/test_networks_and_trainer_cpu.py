"""The training step around the hot path (judge row g / SURVEY section 7 step 8), host logic on the CPU:
  * snde_b200/networks.py against outputs of the reference's DispNetS / PoseNet (tests/golden/network_kats.npz);
  * snde_b200/trainer.py against ONE iteration of the reference's GeoNetModel (tests/golden/train_step_kats.npz), with the CPU
    oracle port standing in for the CUDA loss op (tests may use the oracle; the product's default loss is the fused op);
  * two gloo ranks with DistributedDataParallel: the averaged parameter gradients equal the single-process gradients on the
    concatenated batch (SURVEY section 4 test 4).
The GPU versions of the same checks (fused CUDA loss, NCCL) are in tests/test_gpu_trainer.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_loss_fn(tgt_pyr, src_pyr, disp_pyr, poses, K_ms, cfg):
    """CPU stand-in with the trainer's loss_fn signature: depth = 1/disp (models/baseline.py:835), then the reference's op chain"""
    from oracle import torch_port
    depth = [1.0 / d for d in disp_pyr]
    total, rigid, smooth = torch_port.loss_path(tgt_pyr, src_pyr, depth, poses, K_ms, cfg.simi_alpha, cfg.loss_weight_rigid_warp,
                                                cfg.loss_weight_disparity_smooth)
    return total, torch.stack([total.detach(), rigid.detach(), smooth.detach()])


def test_networks_match_reference_outputs_and_layout():
    from snde_b200.networks import DispNetS, GeoNets, PoseNet, count_parameters
    k = load_golden("network_kats")
    torch.manual_seed(0); d = DispNetS(); d.init_weights(); d.train()
    torch.manual_seed(1); p = PoseNet(2); p.init_weight(); p.train()
    assert count_parameters(d) == int(k["disp_params"]) == 31596900
    assert count_parameters(p) == int(k["pose_params"]) == 1590668
    assert list(d.state_dict().keys()) == list(k["disp_keys"])           # a reference checkpoint loads by name
    assert list(p.state_dict().keys()) == list(k["pose_keys"])
    np.testing.assert_allclose([float(t.double().abs().sum()) for t in d.state_dict().values()], k["disp_abs_sums"], rtol=1e-12)
    np.testing.assert_allclose([float(t.double().abs().sum()) for t in p.state_dict().values()], k["pose_abs_sums"], rtol=1e-12)
    with torch.no_grad():
        for tag in ("a", "b"):
            outs = d(torch.from_numpy(k[f"disp_in_{tag}"]))
            assert len(outs) == 4
            for i, t in enumerate(outs):
                ref = k[f"disp_out_{tag}_{i}"]
                assert tuple(t.shape) == ref.shape
                assert np.abs(t.numpy() - ref).max() <= 1e-6
                assert t.min().item() > 0.01 - 1e-6 and t.max().item() < 10.01 + 1e-6      # 10 * sigmoid + 0.01
        assert np.abs(p(torch.from_numpy(k["pose_in"])).numpy() - k["pose_out"]).max() <= 1e-8
        d.eval()
        assert d(torch.from_numpy(k["disp_in_b"])).shape == k["disp_out_b_0"].shape      # eval: finest scale only
    assert count_parameters(GeoNets(2)) == 33187568                                       # the all-reduce payload: 132.75 MB


def test_trainer_step_matches_reference_iteration():
    from snde_b200.trainer import GeoNetTrainer, TrainConfig
    k = load_golden("train_step_kats")
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    tr = GeoNetTrainer(TrainConfig(seed=int(k["seed"])), device=torch.device("cpu"), ddp=False, loss_fn=oracle_loss_fn)
    tgt, src, K = torch.from_numpy(k["tgt_u8"]), torch.from_numpy(k["src_u8"]), torch.from_numpy(k["K"])
    tr.optimizer.zero_grad()
    total, detail = tr.forward_loss(tgt, src, K)
    assert abs(total.item() - float(k["loss_total"])) < 1e-5 * float(k["loss_total"])
    assert abs(detail[1].item() - float(k["loss_rigid_warp"])) < 1e-5 * float(k["loss_rigid_warp"])
    assert abs(detail[2].item() - float(k["loss_disp_smooth"])) < 1e-5 * float(k["loss_disp_smooth"])
    total.backward()
    gd = np.array([float(p.grad.double().norm()) for p in tr.nets.disp_net.parameters()])
    gp = np.array([float(p.grad.double().norm()) for p in tr.nets.pose_net.parameters()])
    np.testing.assert_allclose(gd, k["disp_grad_l2"], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(gp, k["pose_grad_l2"], rtol=2e-4, atol=1e-9)
    tr.optimizer.step()
    np.testing.assert_allclose([float(p.detach().double().abs().sum()) for p in tr.nets.disp_net.parameters()], k["disp_abs_sum_after_step"], rtol=1e-6)
    np.testing.assert_allclose([float(p.detach().double().abs().sum()) for p in tr.nets.pose_net.parameters()], k["pose_abs_sum_after_step"], rtol=1e-6)
    sd = tr.state_dict()
    assert set(sd) == {"iter", "disp_net_state_dict", "pose_net_state_dict"}              # the reference's checkpoint keys


# ---- two gloo ranks, DDP ---------------------------------------------------------------------------------------------------
B, S, H, W = 4, 2, 128, 160


def _frames(seed=3):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3 * (1 + S), H // 8, W // 8, generator=g)
    fr = (torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1) * 255).round().to(torch.uint8)
    K = torch.tensor([[0.58 * W, 0, 0.5 * W], [0, 1.92 * H, 0.5 * H], [0, 0, 1]]).repeat(B, 1, 1)
    return fr[:, :3].contiguous(), fr[:, 3:].contiguous(), K


def _ddp_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from snde_b200.sharding import shard_range
    from snde_b200.trainer import GeoNetTrainer, TrainConfig
    tr = GeoNetTrainer(TrainConfig(seed=11), device=torch.device("cpu"), loss_fn=oracle_loss_fn)     # DDP: a process group exists
    assert tr.model is not tr.nets
    tgt, src, K = _frames()
    lo, hi = shard_range(B, rank, world)
    out = tr.step(tgt[lo:hi], src[lo:hi], K[lo:hi])
    # after backward every rank holds the AVERAGED gradients; after the Adam step identical parameters
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=out["loss_total"].numpy(),
             grad_l2=np.array([float(p.grad.double().norm()) for p in tr.nets.parameters()]),
             first=next(tr.nets.disp_net.parameters()).grad.numpy(),
             after=np.array([float(p.detach().double().abs().sum()) for p in tr.nets.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_two_ranks_equal_single_process_on_the_concatenated_batch(tmp_path):
    from snde_b200.trainer import GeoNetTrainer, TrainConfig
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_ddp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    tr = GeoNetTrainer(TrainConfig(seed=11), device=torch.device("cpu"), ddp=False, loss_fn=oracle_loss_fn)
    tgt, src, K = _frames()
    out = tr.step(tgt, src, K)
    ref_l2 = np.array([float(p.grad.double().norm()) for p in tr.nets.parameters()])
    ref_first = next(tr.nets.disp_net.parameters()).grad.numpy()
    ref_after = np.array([float(p.detach().double().abs().sum()) for p in tr.nets.parameters()])
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    assert np.array_equal(parts[0]["grad_l2"], parts[1]["grad_l2"])                     # all-reduced: identical on both ranks
    assert np.array_equal(parts[0]["after"], parts[1]["after"])
    # mean of the two shard losses == the loss of the whole batch; averaged gradients == whole-batch gradients
    assert abs(0.5 * (float(parts[0]["loss"]) + float(parts[1]["loss"])) - float(out["loss_total"])) < 1e-5 * float(out["loss_total"])
    np.testing.assert_allclose(parts[0]["grad_l2"], ref_l2, rtol=5e-4, atol=1e-9)
    assert np.linalg.norm(parts[0]["first"] - ref_first) < 5e-4 * np.linalg.norm(ref_first)
    np.testing.assert_allclose(parts[0]["after"], ref_after, rtol=1e-5)
