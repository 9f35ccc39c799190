"""GPU tests of the training step around the fused loss (snde_b200/trainer.py) and of the autograd wrapper's plan cache."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from cases import load_golden, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def fp32_convs():
    """exact fp32 convolutions for comparisons against CPU results (the trainer itself leaves PyTorch's defaults alone)"""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_trainer_step_with_fused_loss_matches_reference_iteration(fp32_convs):
    """ONE iteration of the reference's GeoNetModel (tests/golden/train_step_kats.npz, CPU) against the trainer on the GPU with the
    fused CUDA loss: cuDNN vs CPU convolutions differ in the last bits, so losses to 1e-4 and gradient norms to 2e-3."""
    from snde_b200.trainer import GeoNetTrainer, TrainConfig
    k = load_golden("train_step_kats")
    tr = GeoNetTrainer(TrainConfig(seed=int(k["seed"])), device=torch.device(DEV), ddp=False)
    tgt, src, K = (torch.from_numpy(k[n]).to(DEV) for n in ("tgt_u8", "src_u8", "K"))
    tr.optimizer.zero_grad()
    total, detail = tr.forward_loss(tgt, src, K)
    total.backward()
    torch.cuda.synchronize()
    print(f"[trainer] loss {total.item():.7f} vs reference {float(k['loss_total']):.7f}")
    assert abs(total.item() - float(k["loss_total"])) < 1e-4 * float(k["loss_total"])
    assert abs(detail[1].item() - float(k["loss_rigid_warp"])) < 1e-4 * float(k["loss_rigid_warp"])
    assert abs(detail[2].item() - float(k["loss_disp_smooth"])) < 1e-4 * float(k["loss_disp_smooth"])
    gd = np.array([float(p.grad.double().norm()) for p in tr.nets.disp_net.parameters()])
    gp = np.array([float(p.grad.double().norm()) for p in tr.nets.pose_net.parameters()])
    print(f"[trainer] grad-norm deviation: disp {np.abs(gd / k['disp_grad_l2'] - 1).max():.2e}, pose {np.abs(gp / k['pose_grad_l2'] - 1).max():.2e}")
    np.testing.assert_allclose(gd, k["disp_grad_l2"], rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(gp, k["pose_grad_l2"], rtol=2e-3, atol=1e-9)
    tr.optimizer.step()
    np.testing.assert_allclose([float(p.detach().double().abs().sum()) for p in tr.nets.disp_net.parameters()], k["disp_abs_sum_after_step"], rtol=2e-3)   # (Adam moves every weight by ~lr * sign(g): elements with |g| ~ eps flip with the last bits of cuDNN vs CPU convolutions)


def test_trainer_learns_on_a_fixed_batch_and_reuses_one_plan():
    """fresh network outputs every step must hit the plan cache (keyed on shapes, not addresses) and not pile up memory;
    a few Adam steps on one batch lower the loss"""
    from snde_b200 import loss_functions as LF
    from snde_b200.trainer import GeoNetTrainer, TrainConfig
    k = load_golden("train_step_kats")
    LF._PLAN_CACHE.clear()
    tr = GeoNetTrainer(TrainConfig(seed=3), device=torch.device(DEV), ddp=False)
    tgt, src, K = (torch.from_numpy(k[n]).to(DEV) for n in ("tgt_u8", "src_u8", "K"))
    losses, mem = [], []
    for i in range(12):
        out = tr.step(tgt, src, K)
        losses.append(out["loss_total"].item())
        torch.cuda.synchronize()
        mem.append(torch.cuda.memory_allocated())
    print("[trainer] losses", [round(x, 4) for x in losses])
    assert len(LF._PLAN_CACHE) == 1
    assert mem[-1] <= mem[3] + (1 << 20)                      # steady state: nothing accumulates per step
    assert all(np.isfinite(losses)) and min(losses[6:]) < losses[0]


def test_unit_upstream_and_scaled_backward_agree_and_own_their_gradients():
    from snde_b200.loss_functions import geonet_loss
    from snde_b200.synthetic import make_inputs
    d = make_inputs(B=2, S=2, H=32, W=64, num_scales=3, seed=1, smooth=True).to(DEV)
    def run(scale, unit):
        depth = [t.clone().requires_grad_(True) for t in d.depth_pyr]
        poses = d.poses.clone().requires_grad_(True)
        total, _ = geonet_loss(d.tgt_pyr, d.src_pyr, depth, poses, d.K_ms, unit_upstream=unit)
        (total if scale == 1.0 else scale * total).backward()
        return [t.grad for t in depth], poses.grad
    g1, p1 = run(1.0, True)
    g2, p2 = run(1.0, False)
    g3, p3 = run(2.0, False)
    # a second graph alive at the same time must not overwrite the first one's gradients (each call owns its outputs)
    depth_a = [t.clone().requires_grad_(True) for t in d.depth_pyr]
    ta, _ = geonet_loss(d.tgt_pyr, d.src_pyr, depth_a, d.poses, d.K_ms)
    depth_b = [(t * 1.5).requires_grad_(True) for t in d.depth_pyr]
    tb, _ = geonet_loss(d.tgt_pyr, d.src_pyr, depth_b, d.poses, d.K_ms)
    ta.backward(); tb.backward()
    for a, b, c, e in zip(g1, g2, g3, depth_a):
        assert torch.equal(a, b) and torch.equal(2.0 * a, c) and torch.equal(e.grad, a)
    assert torch.equal(p1, p2) and torch.allclose(2.0 * p1, p3)
    assert not torch.equal(depth_b[0].grad, g1[0])


def test_flow_smooth_loss_intended_semantics():
    """the reference's own body raises; the intended computation is pinned by the reference's smooth_loss
    applied per flow channel (tests/golden/glue_kats.npz: fsl_*)"""
    from snde_b200.loss_functions import flow_smooth_loss
    k = load_golden("glue_kats")
    flow = torch.from_numpy(k["fsl_flow"]).to(DEV).requires_grad_(True)
    img = torch.from_numpy(k["fsl_img"]).to(DEV)
    loss = flow_smooth_loss(flow, img)
    assert abs(loss.item() - float(k["fsl_loss"])) < 1e-5 * float(k["fsl_loss"])
    loss.backward()
    assert rel_l2(flow.grad.cpu().numpy(), k["fsl_gflow"]) < 1e-5
    with pytest.raises(ValueError):
        flow_smooth_loss(flow[:, :1], img)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_ddp_nccl_two_gpus_equal_single_gpu_on_the_concatenated_batch(tmp_path):
    """SURVEY section 4 test 4 on real hardware: 2 ranks x B/2 with NCCL all-reduce == 1 rank x B (gradient norms, parameters after
    one Adam step).  Worker: tests/ddp_worker.py under torchrun."""
    out = str(tmp_path / "ddp.npz")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "ddp_worker.py"), out]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    d = np.load(out)
    print(f"[ddp] loss 2 ranks {float(d['loss_ddp']):.7f} vs 1 rank {float(d['loss_single']):.7f}; max grad-norm deviation "
          f"{np.abs(d['l2_ddp'] / d['l2_single'] - 1).max():.2e}; all-reduced bytes per step {int(d['bytes'])}")
    assert abs(float(d["loss_ddp"]) - float(d["loss_single"])) < 1e-5 * float(d["loss_single"])
    np.testing.assert_allclose(d["l2_ddp"], d["l2_single"], rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(d["after_ddp"], d["after_single"], rtol=2e-3)     # (Adam: ~lr * sign(g) per weight, see above)
    assert int(d["bytes"]) == 4 * 33187568
