"""One training step of the reference's self-supervised trainer around the fused loss -- the caller of the hot path.

Thin re-expression of ``GeoNetModel.training_inside_epoch`` (models/baseline.py:1170-1206) with the loss path replaced by
the fused CUDA op:

    iter_data_preparation   uint8 frames -> [-1,1] floats + both pyramids     one kernel per frame set (snde_frames_to_pyramid)
    build_dispnet           DispNetS on cat(tgt, src_1..S) -> 4 disparities   cuDNN (snde_b200/networks.py)
    build_posenet           PoseNet on cat(tgt, src) channels -> (B,S,6)      cuDNN
    build_rigid_warp_flow + build_losses + the backward of both               ONE fused launch (geonet_loss, disparity=True:
                                                                              depth = 1/disp and its backward are in-kernel)
    optimizer.zero_grad / loss_total.backward / optimizer.step                autograd through the two networks, Adam

Data parallelism: one process per GPU, the batch sharded by sample; every rank takes the mean over
its own shard, ``DistributedDataParallel`` averages the 33,187,568 network gradients (132.75 MB, S = 2) over NCCL while the
backward is still running (bucketed all-reduce).  Equal shards make that average the gradient of the global-batch mean.  There
is no collective inside the loss.

Nothing here is reference code: hyper-parameters are the reference's defaults (models/baseline.py:43-123: Adam lr 2e-4,
betas (0.9, 0.999), simi_alpha 0.85, loss weights 1.0 / 0.5, 4 scales).
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn

from .networks import GeoNets


@dataclasses.dataclass
class TrainConfig:
    num_source: int = 2
    num_scales: int = 4
    simi_alpha: float = 0.85
    loss_weight_rigid_warp: float = 1.0
    loss_weight_disparity_smooth: float = 0.5
    learning_rate: float = 2e-4
    momentum: float = 0.9            # Adam beta1 (the reference calls it momentum)
    beta: float = 0.999
    weight_decay: float = 0.0
    seed: int = 8964


def _default_loss(tgt_pyr, src_pyr, disp_pyr, poses, K_ms, cfg: TrainConfig):
    from .loss_functions import geonet_loss
    return geonet_loss(tgt_pyr, src_pyr, disp_pyr, poses, K_ms, cfg.simi_alpha, cfg.loss_weight_rigid_warp,
                       cfg.loss_weight_disparity_smooth, disparity=True, unit_upstream=True)


class GeoNetTrainer:
    """step(tgt_u8 (B,3,H,W) uint8, src_u8 (B,3*S,H,W) uint8, intrinsics (B,3,3)) -> dict of loss scalars (device tensors).

    ``loss_fn(tgt_pyr, src_pyr, disp_pyr, poses, K_ms, cfg) -> (loss_total, detail)`` is injectable for the host-logic tests
    (a CPU stand-in for the CUDA op); the product uses the fused op.
    With ``torch.distributed`` initialised the networks are wrapped in DDP (device_ids = this rank's GPU, NCCL)."""

    def __init__(self, cfg: TrainConfig = TrainConfig(), device: Optional[torch.device] = None, ddp: Optional[bool] = None,
                 loss_fn: Optional[Callable] = None, bucket_cap_mb: int = 25, channels_last: bool = True):
        import torch.distributed as dist
        self.cfg = cfg
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        torch.manual_seed(cfg.seed)                      # every rank builds identical weights (DDP also broadcasts rank 0's)
        self.nets = GeoNets(cfg.num_source).to(self.device).train()
        if self.device.type == "cuda" and channels_last:
            # NHWC weights/activations: cuDNN's tensor-core convolution kernels take them without layout transposes (measured
            # 35.0 -> 29.5 ms per step at config 2; same arithmetic -- profiles/r2_train_variants.txt)
            self.nets.to(memory_format=torch.channels_last)
        self.loss_fn = loss_fn or _default_loss
        use_ddp = (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) if ddp is None else ddp
        self.model: nn.Module = self.nets
        if use_ddp:
            from torch.nn.parallel import DistributedDataParallel as DDP
            kw = dict(device_ids=[self.device.index]) if self.device.type == "cuda" else {}
            self.model = DDP(self.nets, gradient_as_bucket_view=True, bucket_cap_mb=bucket_cap_mb, **kw)
        self.optimizer = torch.optim.Adam(
            [{"params": self.nets.disp_net.parameters(), "lr": cfg.learning_rate},
             {"params": self.nets.pose_net.parameters(), "lr": cfg.learning_rate}],
            betas=(cfg.momentum, cfg.beta), weight_decay=cfg.weight_decay, **({"fused": True} if self.device.type == "cuda" else {}))
        self.n_iter = 0

    # ---- the pieces of one iteration ---------------------------------------------------------------------------------
    def forward_loss(self, tgt_u8: torch.Tensor, src_u8: torch.Tensor, intrinsics: torch.Tensor):
        """networks forward + fused loss; returns (loss_total, detail).  The networks read level 0 of the very pyramids the
        loss kernel reads (the normalised frames are formed once, on the device)."""
        cfg = self.cfg
        B, S, ns = tgt_u8.shape[0], cfg.num_source, cfg.num_scales
        H, W = tgt_u8.shape[2], tgt_u8.shape[3]
        if tgt_u8.is_cuda:
            from .loss_functions import _frame_pyramids, _intrinsics_pyramid
            tgt_pyr, src_pyr = _frame_pyramids(tgt_u8, src_u8, ns)                # one kernel per frame set
            K_ms = _intrinsics_pyramid(intrinsics, ns)
        else:                                                                     # host-logic tests only (CPU stand-in loss)
            from .synthetic import multi_scale_intrinsics, scale_pyramid
            norm = lambda u: u.float() * (1. / 255.) * 2. - 1.
            tgt_pyr = scale_pyramid(norm(tgt_u8), ns)
            src_pyr = scale_pyramid(torch.cat([norm(src_u8[:, 3 * k:3 * (k + 1)]) for k in range(S)], 0), ns)
            K_ms = multi_scale_intrinsics(intrinsics.float(), ns)
        frames = torch.cat((tgt_pyr[0], src_pyr[0]), 0)                           # (B(1+S),3,H,W): target, then sources source-major
        pose_in = torch.cat((tgt_pyr[0], src_pyr[0].view(S, B, 3, H, W).transpose(0, 1).reshape(B, 3 * S, H, W)), 1)
        disps, poses = self.model(frames, pose_in)
        disp_pyr = [d.squeeze(1) for d in disps[:ns]]
        out = self.loss_fn(tgt_pyr, src_pyr, disp_pyr, poses, K_ms, cfg)
        return out[0], out[1]

    def step(self, tgt_u8: torch.Tensor, src_u8: torch.Tensor, intrinsics: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.optimizer.zero_grad(set_to_none=True)
        total, detail = self.forward_loss(tgt_u8, src_u8, intrinsics)
        total.backward()                                 # DDP: bucketed NCCL all-reduce overlaps this
        self.optimizer.step()
        self.n_iter += 1
        return {"loss_total": detail[0], "loss_rigid_warp": detail[1], "loss_disp_smooth": detail[2]}

    # ---- checkpoint: the reference's dictionary keys (models/baseline.py:1226-1231) ------------------------------------
    def state_dict(self) -> dict:
        return {"iter": self.n_iter, "disp_net_state_dict": self.nets.disp_net.state_dict(),
                "pose_net_state_dict": self.nets.pose_net.state_dict()}

    def load_state_dict(self, sd: dict) -> None:
        self.nets.disp_net.load_state_dict(sd["disp_net_state_dict"])
        self.nets.pose_net.load_state_dict(sd["pose_net_state_dict"])
        self.n_iter = int(sd.get("iter", 0))

    def gradient_bytes(self) -> int:
        return 4 * sum(p.numel() for p in self.nets.parameters())
