"""Oracle of the depth <-> normal consistency path (BASELINE config 3) -- TEST / BASELINE INFRASTRUCTURE ONLY.

PARITY UNPINNED BY THE REFERENCE.  What the reference has is the inference-only GeoNet (Qi et al., CVPR 2018) refinement inside
``NNET.forward`` (models/baseline.py): depth-to-normal (D2N, l.350-446) and normal-to-depth (N2D, l.495-547) over a k x k = 9 x 9
window with dilation 4.  It cannot serve as a parity target as it stands:
  * its patch extraction is scrambled -- ``F.unfold`` is applied to NHWC tensors (l.357, l.396, l.495), so the "patches" mix
    channels and columns (the centre tap is not the pixel's own normal);
  * the batch size is hard-coded to 4, ``torch.det`` makes a GPU -> CPU round trip (l.416), the encoder needs torch.hub;
  * there is NO consistency loss anywhere in the repository and no backward.
This module therefore restates the INTENDED computation: the reference's formulas line by line, with the window taken over the
spatial neighbourhood as in GeoNet (the one-line layout fix: unfold the NCHW tensor), and defines the loss on top of them:

  grid     g_i = (x_i, y_i, 1), x = linspace(-0.6, 0.6, w), y = linspace(-0.4, 0.4, h)              l.308-317 (fixed, not from K)
  window   j in N(i): offsets (dy, dx) in {-(k-1)/2 .. (k-1)/2}^2 * rate, zero outside the image     l.357-358
  D2N      angle_ij = n_j . n_i; valid_ij = angle_ij > thresh (0.95)                                 l.376-377
           P_j = d_j g_j; A = rows P_j of the valid j (others 0); M = A^T A; b = A^T 1               l.392-409
           n'_i = (det M > 1e-5 ? M : I)^-1 b, then L2-normalised                                    l.416-436
  N2D      lower_ij = n_j . g_i; cond_ij = lower_ij > 1e-5; ratio_ij = (n_j . g_j) * (cond ? 1/lower : 1)   l.510-515
           w_ij = cond_ij ? angle_ij : 0; d'_i = clamp(sum_j ratio_ij d_j w_ij / (sum_j w_ij + 1e-5), 0, 10)  l.513-525
  loss     w_d * mean_i |d_i - d'_i|  +  w_n * mean_i (1 - n_i . n'_i)                               (defined here)
Gradient: to the depth map, through d_i and through the d_j inside d'_i (which is linear in them: the weights depend on the
normals only; the clamp passes gradient inside (0, 10)).  Normals are data here (like the images of the photometric loss): the
w_n term is reported, it has no gradient path to depth that this path back-propagates (n' is treated as a constant).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _grid(h, w, xmax, ymax, dtype):
    x = torch.linspace(-xmax, xmax, w, dtype=dtype)
    y = torch.linspace(-ymax, ymax, h, dtype=dtype)
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    return torch.stack([xx, yy, torch.ones_like(xx)], 0)            # (3,h,w): (x, y, 1)


def _patches(t, k, rate):
    """(B,C,h,w) -> (B,C,k*k,h,w): the dilated window of every pixel, zero outside the image (F.unfold on NCHW)"""
    B, C, h, w = t.shape
    pad = (k + (k - 1) * (rate - 1) - 1) // 2
    return F.unfold(t, k, dilation=rate, stride=1, padding=pad).view(B, C, k * k, h, w)


def depth_to_normal(depth, normal, k=9, rate=4, thresh=0.95, xmax=0.6, ymax=0.4):
    """depth (B,h,w), normal (B,3,h,w) -> n' (B,3,h,w)"""
    B, h, w = depth.shape
    g = _grid(h, w, xmax, ymax, depth.dtype).unsqueeze(0)
    pts = g * depth.unsqueeze(1)                                      # P = d * (x, y, 1)
    nj, pj = _patches(normal, k, rate), _patches(pts, k, rate)        # (B,3,K,h,w)
    angle = (nj * normal.unsqueeze(2)).sum(1)                         # (B,K,h,w)
    valid = (angle > thresh).to(depth.dtype).unsqueeze(1)
    a = pj * valid                                                    # rows of A (invalid rows zero)
    M = torch.einsum("bikhw,bjkhw->bhwij", a, a)
    b = a.sum(2).permute(0, 2, 3, 1)                                  # A^T 1
    det = torch.det(M)
    eye = torch.eye(3, dtype=depth.dtype).expand_as(M)
    M = torch.where((det > 1e-5)[..., None, None], M, eye)
    n = torch.linalg.solve(M, b.unsqueeze(-1)).squeeze(-1)            # inverse(M) @ b
    return F.normalize(n, p=2, dim=-1).permute(0, 3, 1, 2)


def normal_to_depth(depth, normal, k=9, rate=4, xmax=0.6, ymax=0.4, clamp_max=10.0, return_parts=False):
    """depth (B,h,w), normal (B,3,h,w) -> d' (B,h,w)"""
    B, h, w = depth.shape
    g = _grid(h, w, xmax, ymax, depth.dtype).unsqueeze(0)
    nj = _patches(normal, k, rate)                                    # (B,3,K,h,w)
    gj = _patches(g.expand(B, -1, -1, -1), k, rate)
    dj = _patches(depth.unsqueeze(1), k, rate)[:, 0]                  # (B,K,h,w)
    angle = (nj * normal.unsqueeze(2)).sum(1)
    lower = (nj * g.unsqueeze(2)).sum(1)                              # n_j . g_i
    cond = lower > 1e-5
    inv = torch.where(cond, 1.0 / lower, torch.ones_like(lower))
    upper = (nj * gj).sum(1)                                          # n_j . g_j
    wgt = torch.where(cond, angle, torch.zeros_like(angle))
    wsum = wgt.sum(1) + 1e-5
    raw = (inv * upper * dj * wgt).sum(1) / wsum
    out = raw.clamp(0, clamp_max)
    return (out, raw) if return_parts else out


def consistency_loss(depth, normal, k=9, rate=4, thresh=0.95, xmax=0.6, ymax=0.4, w_depth=1.0, w_normal=1.0):
    """-> (loss, depth term, normal term).  Autograd through this gives the depth gradient of the spec above when `normal`
    does not require grad and n' is detached."""
    d2 = normal_to_depth(depth, normal, k, rate, xmax, ymax)
    with torch.no_grad():
        n2 = depth_to_normal(depth, normal, k, rate, thresh, xmax, ymax)
    ld = (depth - d2).abs().mean()
    ln = (1.0 - (normal * n2).sum(1)).mean()
    return w_depth * ld + w_normal * ln, ld, ln


def synthetic_scene(B, h, w, seed=0, dtype=torch.float32):
    """Depth of a few tilted planes + ripples, normals = the planes' true normals with a little noise: most neighbours pass the
    0.95 angle test, plane borders and the noise exercise the invalid branches."""
    gen = torch.Generator().manual_seed(seed)
    g = _grid(h, w, 0.6, 0.4, dtype)
    depth = torch.empty(B, h, w, dtype=dtype)
    normal = torch.empty(B, 3, h, w, dtype=dtype)
    for b in range(B):
        n_planes = 3
        best = torch.full((h, w), 1e9, dtype=dtype)
        nb = torch.zeros(3, h, w, dtype=dtype)
        for _ in range(n_planes):
            n = torch.tensor([0.0, 0.0, -1.0], dtype=dtype) + 0.5 * (torch.rand(3, generator=gen, dtype=dtype) - 0.5)
            n = n / n.norm()
            c = -(1.0 + 2.0 * torch.rand((), generator=gen, dtype=dtype))      # plane n . P = c (c < 0 with n_z < 0 -> positive depth)
            z = c / (n.view(3, 1, 1) * g).sum(0)
            z = torch.where(z > 0.2, z, torch.full_like(z, 1e9))
            take = z < best
            best = torch.where(take, z, best)
            nb = torch.where(take.unsqueeze(0), (-n).view(3, 1, 1).expand(3, h, w), nb)
        best = torch.where(best > 1e8, torch.full_like(best, 3.0), best)
        depth[b] = best * (1.0 + 0.01 * torch.rand(h, w, generator=gen, dtype=dtype))
        nn = nb + 0.05 * torch.randn(3, h, w, generator=gen, dtype=dtype)
        normal[b] = nn / nn.norm(dim=0, keepdim=True).clamp_min(1e-6)
    return depth, normal
