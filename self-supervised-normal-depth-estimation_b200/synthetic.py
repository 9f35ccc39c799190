"""Seeded synthetic KITTI-/NYU-shaped inputs for the loss path.

Everything is drawn from a CPU ``torch.Generator`` so CPU and GPU runs see identical
bits.  Draw order (fixed): tgt, src (source-major), per-scale disparity, poses.

``scale_pyramid`` / ``multi_scale_intrinsics`` restate the reference's caller-side prep
(``utils/utils_edited.py:89-112`` and ``:149-169``) as torch glue; they are inputs to the
hot path, not part of it.
"""
from __future__ import annotations

import dataclasses
from typing import List

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class LossInputs:
    """Inputs of one loss evaluation. Images are planar (n,3,h,w); the reference's
    logical NHWC views are ``x.permute(0,2,3,1)``."""
    tgt_pyr: List[torch.Tensor]      # scales x (B,3,h,w)
    src_pyr: List[torch.Tensor]      # scales x (S*B,3,h,w), source-major n = src*B + b
    depth_pyr: List[torch.Tensor]    # scales x (B*(1+S),h,w): target depths then source depths
    poses: torch.Tensor              # (B,S,6)  [tx,ty,tz,rx,ry,rz]
    K: torch.Tensor                  # (B,3,3)
    K_ms: torch.Tensor               # (B,scales,3,3)

    @property
    def B(self): return self.poses.shape[0]

    @property
    def S(self): return self.poses.shape[1]

    @property
    def num_scales(self): return len(self.tgt_pyr)

    def warped_pixels(self) -> int:
        """2 * N * sum_s h_s*w_s  (both directions)."""
        return 2 * self.B * self.S * sum(t.shape[2] * t.shape[3] for t in self.tgt_pyr)

    def to(self, device):
        f = lambda t: t.to(device)
        return LossInputs([f(t) for t in self.tgt_pyr], [f(t) for t in self.src_pyr],
                          [f(t) for t in self.depth_pyr], f(self.poses), f(self.K), f(self.K_ms))


def scale_pyramid(img: torch.Tensor, num_scales: int) -> List[torch.Tensor]:
    """(b,3,H,W) -> list of planar area-downsampled levels, sizes int(H/2^s) x int(W/2^s)."""
    _, _, h, w = img.shape
    out = [img]
    for i in range(1, num_scales):
        r = 2 ** i
        out.append(F.interpolate(img, size=(int(h / r), int(w / r)), mode="area"))
    return out


def multi_scale_intrinsics(K: torch.Tensor, num_scales: int) -> torch.Tensor:
    """(B,3,3) -> (B,scales,3,3): fx,fy,cx,cy / 2^s, zeros and a one elsewhere."""
    out = torch.zeros(K.shape[0], num_scales, 3, 3, dtype=K.dtype, device=K.device)
    for s in range(num_scales):
        out[:, s, 0, 0] = K[:, 0, 0] / (2 ** s)
        out[:, s, 1, 1] = K[:, 1, 1] / (2 ** s)
        out[:, s, 0, 2] = K[:, 0, 2] / (2 ** s)
        out[:, s, 1, 2] = K[:, 1, 2] / (2 ** s)
        out[:, s, 2, 2] = 1.0
    return out


def _smooth_images(g, n, H, W):
    low = torch.rand(n, 3, max(H // 8, 2), max(W // 8, 2), generator=g)
    up = F.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
    return up.clamp(0, 1) * 2 - 1


def make_inputs(B=4, S=2, H=128, W=416, num_scales=4, seed=0, smooth=False, motion="small",
                intrinsics="kitti") -> LossInputs:
    """motion: 'small' (PoseNet scale, randn*0.01) or 'large' (t~0.1, r~0.02: most pixels leave the view).
    smooth=True gives band-limited frames / depth (the parity variant)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if smooth:
        tgt = _smooth_images(g, B, H, W)
        src = _smooth_images(g, B * S, H, W)
    else:
        tgt = torch.rand(B, 3, H, W, generator=g) * 2 - 1
        src = torch.rand(B * S, 3, H, W, generator=g) * 2 - 1
    if intrinsics == "kitti":
        K1 = torch.tensor([[0.58 * W, 0, 0.5 * W], [0, 1.92 * H, 0.5 * H], [0, 0, 1]], dtype=torch.float32)
    else:  # nyu
        K1 = torch.tensor([[518.9 * W / 640, 0, 325.6 * W / 640], [0, 519.5 * H / 480, 253.7 * H / 480], [0, 0, 1]],
                          dtype=torch.float32)
    K = K1.repeat(B, 1, 1)
    M = B * (1 + S)
    depth_pyr = []
    for s in range(num_scales):
        h, w = int(H / 2 ** s), int(W / 2 ** s)
        if smooth:
            low = torch.rand(M, 1, max(h // 16, 2), max(w // 16, 2), generator=g)
            u = F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)[:, 0]
        else:
            u = torch.rand(M, h, w, generator=g)
        disp = (u * 0.9 + 0.05) * 10 + 0.01
        depth_pyr.append(1.0 / disp)
    if motion == "small":
        poses = torch.randn(B, S, 6, generator=g) * 0.01
    else:
        poses = torch.randn(B, S, 6, generator=g) * torch.tensor([0.1, 0.1, 0.1, 0.02, 0.02, 0.02])
    return LossInputs(scale_pyramid(tgt, num_scales), scale_pyramid(src, num_scales), depth_pyr, poses, K,
                      multi_scale_intrinsics(K, num_scales))
