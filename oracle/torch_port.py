"""torch (CPU, multi-threaded) restatement of the reference loss path -- TEST / BASELINE INFRASTRUCTURE ONLY.

Purpose: the CPU baseline.  The reference runs this path as a long chain of ATen ops on the host cores; this
file restates that chain op for op (same tensors materialised, same autograd backward), so timing it on the
GPU box's host cores stands in for timing the reference there (`/root/reference` does not travel).  It is
validated against the golden vectors in tests/test_oracle_vs_golden.py.  Only bench.py's cpu_baseline /
--impl reference legs and tests import it.

reference: utils/utils_edited.py:121-141 (DSSIM), :171-289 (pose), :216-266, :291-362 (rigid flow),
:386-486 (flow_warp / bilinear_sampler); models/loss_functions.py:4-24; composition models/baseline.py:895-970,
1125-1161.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _grid(n, dev):
    return (torch.linspace(-1, 1, n, device=dev) + 1) * 0.5 * (n - 1)


def pose_vec2mat(vec):
    z, y, x = vec[:, 5], vec[:, 4], vec[:, 3]
    one, zero = torch.ones_like(z), torch.zeros_like(z)
    Rz = torch.stack([torch.cos(z), -torch.sin(z), zero, torch.sin(z), torch.cos(z), zero, zero, zero, one], 1).view(-1, 3, 3)
    Ry = torch.stack([torch.cos(y), zero, torch.sin(y), zero, one, zero, -torch.sin(y), zero, torch.cos(y)], 1).view(-1, 3, 3)
    Rx = torch.stack([one, zero, zero, zero, torch.cos(x), -torch.sin(x), zero, torch.sin(x), torch.cos(x)], 1).view(-1, 3, 3)
    R = torch.matmul(torch.matmul(Rx, Ry), Rz)
    T = torch.cat([R, vec[:, :3].unsqueeze(2)], 2)
    last = torch.tensor([0., 0., 0., 1.], device=vec.device).view(1, 1, 4).repeat(vec.shape[0], 1, 1)
    return torch.cat([T, last], 1)


def compute_rigid_flow(pose, depth, K, reverse):
    b, h, w = depth.shape
    dev = depth.device
    T = pose_vec2mat(pose)
    if reverse:
        T = torch.inverse(T)
    xt = _grid(w, dev).view(1, w).expand(h, w)
    yt = _grid(h, dev).view(h, 1).expand(h, w)
    pix = torch.stack([xt, yt, torch.ones_like(xt)], 0).unsqueeze(0).expand(b, 3, h, w)
    cam = torch.matmul(torch.inverse(K), pix.reshape(b, 3, -1)) * depth.reshape(b, 1, -1)
    cam = torch.cat([cam, torch.ones(b, 1, h * w, device=dev)], 1)
    K4 = torch.cat([torch.cat([K, torch.zeros(b, 3, 1, device=dev)], 2),
                    torch.tensor([0., 0., 0., 1.], device=dev).view(1, 1, 4).repeat(b, 1, 1)], 1)
    u = torch.matmul(torch.matmul(K4, T), cam)
    z = u[:, 2:3] + 1e-10
    src = torch.cat([u[:, :1] / z, u[:, 1:2] / z], 1).view(b, 2, h, w).permute(0, 2, 3, 1)
    return src - pix[:, :2].permute(0, 2, 3, 1)


def bilinear_sampler(imgs, coords):
    n, hs, ws, c = imgs.shape
    x, y = coords[..., 0:1], coords[..., 1:2]
    x0, y0 = torch.floor(x), torch.floor(y)
    x1, y1 = x0 + 1, y0 + 1
    x0s, x1s = x0.clamp(0, ws - 1), x1.clamp(0, ws - 1)
    y0s, y1s = y0.clamp(0, hs - 1), y1.clamp(0, hs - 1)
    wx0, wx1, wy0, wy1 = x1s - x, x - x0s, y1s - y, y - y0s
    base = (torch.arange(n, device=imgs.device) * (hs * ws)).view(n, 1, 1, 1)
    flat = imgs.contiguous().view(-1, c)

    def take(xs, ys):
        idx = (base + ys.long() * ws + xs.long()).reshape(-1)      # integer indices (the reference's float32 ones overflow)
        return torch.index_select(flat, 0, idx).view(n, coords.shape[1], coords.shape[2], c)

    i00, i01, i10, i11 = take(x0s, y0s), take(x0s, y1s), take(x1s, y0s), take(x1s, y1s)
    return (wx0 * wy0) * i00 + (wx0 * wy1) * i01 + (wx1 * wy0) * i10 + (wx1 * wy1) * i11


def flow_warp(img, flow):
    n, h, w, _ = img.shape
    grid = torch.stack([_grid(w, img.device).view(1, w).expand(h, w), _grid(h, img.device).view(h, 1).expand(h, w)], -1)
    return bilinear_sampler(img, grid.unsqueeze(0) + flow)


def dssim(x, y):
    pool = lambda t: F.avg_pool2d(t, 3, 1, 1)
    x, y = x.permute(0, 3, 1, 2), y.permute(0, 3, 1, 2)
    mx, my = pool(x), pool(y)
    sx, sy, sxy = pool(x ** 2) - mx ** 2, pool(y ** 2) - my ** 2, pool(x * y) - mx * my
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = ((2 * mx * my + C1) * (2 * sxy + C2)) / ((mx ** 2 + my ** 2 + C1) * (sx + sy + C2))
    return torch.clamp((1 - ssim.permute(0, 2, 3, 1)) / 2, 0, 1)


def image_similarity(alpha, x, y):
    return alpha * dssim(x, y) + (1 - alpha) * torch.abs(x - y)


def smooth_loss(depth, image):
    gx = lambda t: t[:, :, :-1, :] - t[:, :, 1:, :]
    gy = lambda t: t[:, :-1, :, :] - t[:, 1:, :, :]
    wx = torch.exp(-torch.mean(torch.abs(gx(image)), 3, True))
    wy = torch.exp(-torch.mean(torch.abs(gy(image)), 3, True))
    return torch.mean(torch.abs(gx(depth) * wx)) + torch.mean(torch.abs(gy(depth) * wy))


def frames_to_pyramid(frames_u8, num_scales):
    """models/baseline.py:761-767 (uint8 -> float, *= 1/255, * 2 - 1) + utils/utils_edited.py:89-112 (area pyramid), planar levels"""
    x = frames_u8.float()
    x *= 1. / 255.
    x = x * 2. - 1.
    h, w = x.shape[2:]
    return [x] + [F.interpolate(x, size=(int(h / 2 ** s), int(w / 2 ** s)), mode="area") for s in range(1, num_scales)]


def loss_path(tgt_pyr, src_pyr, depth_pyr, poses, K_ms, alpha=0.85, w_rigid=1.0, w_smooth=0.5):
    """tgt_pyr[s] (B,3,h,w), src_pyr[s] (S*B,3,h,w) planar; depth_pyr[s] (B(1+S),h,w); poses (B,S,6); K_ms (B,scales,3,3)."""
    B, S = poses.shape[:2]
    rigid, smooth = 0, 0
    for s in range(len(tgt_pyr)):
        tgt, src = tgt_pyr[s].permute(0, 2, 3, 1), src_pyr[s].permute(0, 2, 3, 1)
        tile = tgt.repeat(S, 1, 1, 1)
        ff = torch.cat([compute_rigid_flow(poses[:, k], depth_pyr[s][:B], K_ms[:, s], False) for k in range(S)], 0)
        fb = torch.cat([compute_rigid_flow(poses[:, k], depth_pyr[s][B * (k + 1):B * (k + 2)], K_ms[:, s], True) for k in range(S)], 0)
        fe = image_similarity(alpha, tile, flow_warp(src, ff))
        be = image_similarity(alpha, src, flow_warp(tile, fb))
        rigid = rigid + w_rigid * S / 2 * (fe.mean() + be.mean())
        smooth = smooth + w_smooth / 2 ** s * smooth_loss(depth_pyr[s].unsqueeze(3), torch.cat((tgt, src), 0))
    return rigid + smooth, rigid, smooth


def timed_fwd_bwd(inp, iters=3, warmup=1):
    """seconds per fwd+bwd (median) of the loss path on the host cores. inp: snde_b200.synthetic.LossInputs on CPU."""
    import time
    depth = [d.clone().requires_grad_(True) for d in inp.depth_pyr]
    poses = inp.poses.clone().requires_grad_(True)
    ts = []
    for i in range(warmup + iters):
        for t in depth + [poses]:
            t.grad = None
        t0 = time.perf_counter()
        total, _, _ = loss_path(inp.tgt_pyr, inp.src_pyr, depth, poses, inp.K_ms)
        total.backward()
        t1 = time.perf_counter()
        if i >= warmup:
            ts.append(t1 - t0)
    ts.sort()
    return ts[len(ts) // 2], float(total.item())
