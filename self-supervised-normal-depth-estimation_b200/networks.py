"""The two producers of the loss path's inputs: DispNetS (disparity pyramid) and PoseNet (relative poses).

They are NOT part of the accelerated path (plain convolutions, they stay in cuDNN); they exist here
so that the training step the metric names -- networks forward, fused loss, backward, gradient all-reduce, Adam -- can be
run and measured (snde_b200/trainer.py, bench.py --workload train_*).  Architectures follow the reference
(models/DispNetS.py:1-136, models/PoseNet.py:1-56): no normalisation layers, ReLU everywhere, disparity heads
``alpha * sigmoid(.) + beta`` with alpha = 10, beta = 0.01, poses scaled by 0.01.  Sub-module names and the layout of the
``nn.Sequential`` blocks match the reference's, so a reference checkpoint's ``disp_net_state_dict`` / ``pose_net_state_dict``
(models/baseline.py:1226-1231) loads with ``load_state_dict``.  Parameter counts: 31,596,900 and 1,590,668 (S = 2).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _relu_convs(specs) -> nn.Sequential:
    """specs: (kind, cin, cout, k, stride) -> conv, ReLU, conv, ReLU, ... ('T' = transposed, doubles the size)"""
    layers: List[nn.Module] = []
    for kind, cin, cout, k, stride in specs:
        if kind == "T":
            layers.append(nn.ConvTranspose2d(cin, cout, k, stride=stride, padding=(k - 1) // 2, output_padding=stride - 1))
        else:
            layers.append(nn.Conv2d(cin, cout, k, stride=stride, padding=(k - 1) // 2))
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


def _xavier(net: nn.Module, kinds) -> None:
    for m in net.modules():
        if isinstance(m, kinds):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)


def _fit(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """top-left crop to the spatial size of `like` (odd sizes make the decoder one pixel larger than the skip)"""
    return t[:, :, :like.shape[2], :like.shape[3]]


class DispNetS(nn.Module):
    """7-level encoder (stride 2 each), 7-level decoder with skips, disparity heads at the 4 finest decoder levels.
    forward: (n,3,H,W) in [-1,1] -> training: (disp1 (n,1,H,W), disp2 (H/2), disp3 (H/4), disp4 (H/8)); eval: disp1."""

    ENC = [32, 64, 128, 256, 512, 512, 512]
    DEC = [512, 512, 256, 128, 64, 32, 16]
    ENC_K = [7, 5, 3, 3, 3, 3, 3]

    def __init__(self, alpha: float = 10, beta: float = 0.01):
        super().__init__()
        self.alpha, self.beta = alpha, beta
        cin = 3
        for i, (c, k) in enumerate(zip(self.ENC, self.ENC_K), start=1):
            setattr(self, f"conv{i}", _relu_convs([("C", cin, c, k, 2), ("C", c, c, k, 1)]))
            cin = c
        # (registration order = the reference's, so that seeded initialisation draws the same weights: encoder, upconv7..1,
        #  iconv7..1, heads 4..1)
        for j, c in enumerate(self.DEC):               # upconv7 .. upconv1
            setattr(self, f"upconv{7 - j}", _relu_convs([("T", cin, c, 3, 2)]))
            cin = c
        for j, c in enumerate(self.DEC):               # iconv7 .. iconv1
            lvl = 7 - j
            skip = self.ENC[lvl - 2] if lvl >= 2 else 0
            extra = 1 if lvl <= 3 else 0               # the upsampled coarser disparity joins the three finest levels
            setattr(self, f"iconv{lvl}", _relu_convs([("C", c + skip + extra, c, 3, 1)]))
        for lvl in (4, 3, 2, 1):
            setattr(self, f"predict_disp{lvl}", nn.Sequential(nn.Conv2d(self.DEC[7 - lvl], 1, 3, padding=1), nn.Sigmoid()))

    def init_weights(self) -> None:
        _xavier(self, (nn.Conv2d, nn.ConvTranspose2d))

    def _head(self, lvl: int, feat: torch.Tensor) -> torch.Tensor:
        return self.alpha * getattr(self, f"predict_disp{lvl}")(feat) + self.beta

    def forward(self, x: torch.Tensor):
        skips = []
        t = x
        for i in range(1, 8):
            t = getattr(self, f"conv{i}")(t)
            skips.append(t)
        disps = {}
        prev_disp = None
        for lvl in range(7, 0, -1):
            like = skips[lvl - 2] if lvl >= 2 else x
            parts = [_fit(getattr(self, f"upconv{lvl}")(t), like)]
            if lvl >= 2:
                parts.append(like)
            if lvl <= 3:
                parts.append(_fit(F.interpolate(prev_disp, scale_factor=2, mode="bilinear", align_corners=False), like))
            t = getattr(self, f"iconv{lvl}")(torch.cat(parts, 1))
            if lvl <= 4:
                prev_disp = disps[lvl] = self._head(lvl, t)
        if self.training:
            return disps[1], disps[2], disps[3], disps[4]
        return disps[1]


class PoseNet(nn.Module):
    """7 stride-2 convolutions + a 1x1 head, spatial mean, x0.01.  forward: (B, 3*(1+S), H, W) -> (B, S, 6)
    [tx, ty, tz, rx, ry, rz] per source view."""

    WIDTHS = [16, 32, 64, 128, 256, 256, 256]
    KERNELS = [7, 5, 3, 3, 3, 3, 3]

    def __init__(self, num_source: int):
        super().__init__()
        self.num_source = num_source
        cin = 3 * (1 + num_source)
        for i, (c, k) in enumerate(zip(self.WIDTHS, self.KERNELS), start=1):
            setattr(self, f"conv{i}", _relu_convs([("C", cin, c, k, 2)]))
            cin = c
        self.pred_poses = nn.Conv2d(cin, 6 * num_source, kernel_size=1)

    def init_weight(self) -> None:
        _xavier(self, (nn.Conv2d,))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for i in range(1, 8):
            x = getattr(self, f"conv{i}")(x)
        p = self.pred_poses(x).mean(dim=(2, 3))
        return 0.01 * p.view(p.shape[0], self.num_source, 6)


class GeoNets(nn.Module):
    """Both networks behind one forward, so that ONE DistributedDataParallel wrapper buckets and all-reduces all
    33,187,568 gradients (S = 2) of a step.  forward(frames (B*(1+S),3,H,W) = target frames then the sources source-major,
    pose_in (B,3*(1+S),H,W)) -> (disparities: 4 x (B*(1+S),1,h_s,w_s), poses (B,S,6))
    -- what build_dispnet / build_posenet produce (models/baseline.py:800-865)."""

    def __init__(self, num_source: int):
        super().__init__()
        self.disp_net = DispNetS()
        self.pose_net = PoseNet(num_source)
        self.disp_net.init_weights()
        self.pose_net.init_weight()

    def forward(self, frames: torch.Tensor, pose_in: torch.Tensor) -> Tuple[Tuple[torch.Tensor, ...], torch.Tensor]:
        return self.disp_net(frames), self.pose_net(pose_in)


def count_parameters(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
