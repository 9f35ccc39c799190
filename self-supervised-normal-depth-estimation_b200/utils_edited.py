"""Drop-in names of the reference's ``utils/utils_edited.py`` hot-path functions, backed by libsnde.so.

Signatures, argument meaning, layouts (logical NHWC images, (B,h,w) depth, (B,h,w,2) flow) and
return shapes follow the reference (file:line cited per function).  Differences, all deliberate:
  * gather indices are integers, so N*h*w > 2^24 samples the right pixels (the reference's float32
    indices do not);
  * the module-level ``device`` global (reference :11-12, ``cuda:1``) is gone: ops run on the
    device of their inputs and on the current stream;
  * CUDA only -- there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


# ---- caller-side prep kept as torch glue (not on the per-pixel path) -----------------------------
def scale_pyramid(img, num_scales):
    """reference :89-112 -- NHWC *views* of area-downsampled levels."""
    if img is None:
        return None
    b, ch, h, w = img.shape
    out = [img.permute(0, 2, 3, 1)]
    for i in range(num_scales - 1):
        r = 2 ** (i + 1)
        out.append(F.interpolate(img, size=(int(h / r), int(w / r)), mode="area").permute(0, 2, 3, 1))
    return out


def compute_multi_scale_intrinsics(intrinsics, num_scales):
    """reference :149-169 -> (B, num_scales, 3, 3): fx, fy, cx, cy divided by 2^s, zeros and a one elsewhere."""
    div = (2.0 ** torch.arange(num_scales, device=intrinsics.device, dtype=intrinsics.dtype)).view(1, num_scales)
    out = torch.zeros(intrinsics.shape[0], num_scales, 3, 3, dtype=intrinsics.dtype, device=intrinsics.device)
    out[:, :, 0, 0] = intrinsics[:, 0, 0, None] / div
    out[:, :, 1, 1] = intrinsics[:, 1, 1, None] / div
    out[:, :, 0, 2] = intrinsics[:, 0, 2, None] / div
    out[:, :, 1, 2] = intrinsics[:, 1, 2, None] / div
    out[:, :, 2, 2] = 1.0
    return out


def gradient_x(img):
    """reference :143-144"""
    return img[:, :, :-1, :] - img[:, :, 1:, :]


def gradient_y(img):
    """reference :146-147"""
    return img[:, :-1, :, :] - img[:, 1:, :, :]


def euler2mat(z, y, x):
    """reference :171-214 -- R = Rx(x) @ Ry(y) @ Rz(z), shapes (B,) -> (B,3,3). torch glue."""
    one, zero = torch.ones_like(z), torch.zeros_like(z)
    cz, sz, cy, sy, cx, sx = torch.cos(z), torch.sin(z), torch.cos(y), torch.sin(y), torch.cos(x), torch.sin(x)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).view(-1, 3, 3)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).view(-1, 3, 3)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).view(-1, 3, 3)
    return Rx @ Ry @ Rz


def pose_vec2mat(vec):
    """reference :268-289 -- (B,6) [tx,ty,tz,rx,ry,rz] -> (B,4,4). torch glue."""
    R = euler2mat(vec[:, 5], vec[:, 4], vec[:, 3])
    top = torch.cat([R, vec[:, :3].unsqueeze(2)], 2)
    bot = torch.tensor([0., 0., 0., 1.], device=vec.device, dtype=vec.dtype).view(1, 1, 4).expand(vec.shape[0], 1, 4)
    return torch.cat([top, bot], 1)


def meshgrid(batch, height, width, is_homogeneous=True, device=None):
    """reference :291-328 -- (batch, 3|2, h, w) expanded view built from the bit-identical 1-D tables."""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    xt = ops.grid_table(width, device).view(1, width).expand(height, width)
    yt = ops.grid_table(height, device).view(height, 1).expand(height, width)
    parts = [xt, yt] + ([torch.ones_like(xt)] if is_homogeneous else [])
    return torch.stack(parts, 0).unsqueeze(0).expand(batch, -1, height, width)


def pixel2cam(depth, pixel_coords, intrinsics, is_homogeneous=True):
    """reference :216-239. torch glue (the fused kernels do this per pixel in registers)."""
    b, h, w = depth.shape
    cam = torch.matmul(torch.inverse(intrinsics), pixel_coords.reshape(b, 3, -1)) * depth.reshape(b, 1, -1)
    if is_homogeneous:
        cam = torch.cat([cam, torch.ones(b, 1, h * w, device=cam.device, dtype=cam.dtype)], 1)
    return cam.view(b, -1, h, w)


def cam2pixel(cam_coords, proj):
    """reference :241-266. torch glue."""
    b, _, h, w = cam_coords.shape
    u = torch.matmul(proj, cam_coords.reshape(b, 4, h * w))
    z = u[:, 2:3] + 1e-10
    return torch.cat([u[:, :1] / z, u[:, 1:2] / z], 1).view(b, 2, h, w).permute(0, 2, 3, 1)


# ---- the per-pixel path: CUDA kernels -------------------------------------------------------------
class _RigidFlow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, depth, intrinsics, reverse_pose):
        pose, depth, intrinsics = pose.detach(), depth.detach(), intrinsics.detach()
        P, Kinv = ops.pose_matrices(pose, intrinsics, reverse_pose)
        flow = ops.rigid_flow_fwd(depth, Kinv, P)
        ctx.save_for_backward(pose, depth, intrinsics, P, Kinv)
        ctx.reverse = bool(reverse_pose)
        return flow

    @staticmethod
    def backward(ctx, gflow):
        pose, depth, K, P, Kinv = ctx.saved_tensors
        want_K = ctx.needs_input_grad[2]
        gdepth, gP, gKinv = ops.rigid_flow_bwd(depth, Kinv, P, gflow, want_K)
        gpose, gK = ops.pose_matrices_bwd(pose, K, ctx.reverse, gP, gKinv, want_K)
        return (gpose if ctx.needs_input_grad[0] else None, gdepth if ctx.needs_input_grad[1] else None, gK, None)


def compute_rigid_flow(pose, depth, intrinsics, reverse_pose):
    """reference :331-362 -- pose (B,6), depth (B,h,w), intrinsics (B,3,3) -> rigid flow (B,h,w,2)."""
    return _RigidFlow.apply(pose, depth, intrinsics, bool(reverse_pose))


class _Sampler(torch.autograd.Function):
    """want_valid: also return the (N,h,w) validity mask from the same launch (non-differentiable)."""

    @staticmethod
    def forward(ctx, imgs, coords, add_grid, padding_mode, want_valid=False):
        imgs_d, coords_d = imgs.detach(), coords.detach()
        out, valid = ops.sampler_fwd(imgs_d, coords_d, add_grid, padding_mode, want_valid)
        ctx.save_for_backward(imgs_d, coords_d)
        ctx.cfg = (add_grid, padding_mode)
        if want_valid:
            ctx.mark_non_differentiable(valid)
            return out, valid
        return out

    @staticmethod
    def backward(ctx, gout, _gvalid=None):
        imgs, coords = ctx.saved_tensors
        gcoords, gimg = ops.sampler_bwd(imgs, coords, ctx.cfg[0], ctx.cfg[1], gout, ctx.needs_input_grad[0])
        return gimg, (gcoords if ctx.needs_input_grad[1] else None), None, None, None


def bilinear_sampler(imgs, coords, padding_mode="zeros"):
    """reference :402-486 -- imgs (N,hs,ws,C), coords (N,ht,wt,2) absolute pixels -> (N,ht,wt,C).
    'zeros' is the reference's semantics (value 0 outside [0,ws-1) x [0,hs-1)); 'border' is an extension."""
    return _Sampler.apply(imgs.float(), coords.float(), False, padding_mode)


def flow_warp(src_img, flow, padding_mode="zeros"):
    """reference :386-399 -- bilinear_sampler(src_img, meshgrid + flow); the grid add happens in-kernel."""
    return _Sampler.apply(src_img.float(), flow.float(), True, padding_mode)


def DSSIM(x, y):
    """reference :121-141 -- clamp((1 - SSIM_3x3(x, y)) / 2, 0, 1) on (N,h,w,3) tensors."""
    from .loss_functions import image_similarity
    return image_similarity(1.0, x, y)
