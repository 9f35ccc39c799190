"""Drop-in names of the reference's ``models/loss_functions.py`` (image_similarity, smooth_loss,
flow_smooth_loss) plus the north-star compositions (inverse_warp, photometric_reconstruction_loss)
and the fully fused step loss (geonet_loss), all backed by libsnde.so.  CUDA only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .utils_edited import compute_rigid_flow, flow_warp, _Sampler  # noqa: F401  (re-exported like the reference's import *)


class _ImageSimilarity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, x, y):
        xd, yd = x.detach(), y.detach()
        ctx.save_for_backward(xd, yd)
        ctx.alpha = float(alpha)
        return ops.similarity_fwd(alpha, xd, yd)

    @staticmethod
    def backward(ctx, gerr):
        x, y = ctx.saved_tensors
        gx, gy = ops.similarity_bwd(ctx.alpha, x, y, gerr, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return None, gx, gy


def image_similarity(alpha, x, y):
    """reference models/loss_functions.py:4-6 -- alpha*DSSIM(x,y) + (1-alpha)*|x-y| on (N,h,w,3) -> (N,h,w,3)."""
    return _ImageSimilarity.apply(float(alpha), x.float(), y.float())


class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, image):
        d = depth.detach()
        if d.dim() == 4 and d.shape[3] != 1:
            raise ValueError(f"smooth_loss: depth must be (M,h,w,1) or (M,h,w), got {tuple(d.shape)}")
        if tuple(d.shape[:3]) != tuple(image.shape[:3]):
            raise ValueError(f"smooth_loss: depth {tuple(d.shape)} and image {tuple(image.shape)} disagree")
        d3 = d.reshape(d.shape[0], d.shape[1], d.shape[2])
        ctx.save_for_backward(d3, image.detach())
        ctx.shape = depth.shape
        return ops.smooth_fwd(d3, image.detach())

    @staticmethod
    def backward(ctx, g):
        d3, image = ctx.saved_tensors
        return ops.smooth_bwd(d3, image, g).view(ctx.shape), None


def smooth_loss(depth, image):
    """reference models/loss_functions.py:8-24 -- depth (M,h,w,1), image (M,h,w,3) -> scalar
    mean|dx(depth)*exp(-mean_c|dx(image)|)| + same in y.  No gradient to the image (as in the reference)."""
    return _SmoothLoss.apply(depth.float(), image.float())


def flow_smooth_loss(flow, img):
    """reference models/loss_functions.py:26-32 -- mean over the two flow channels of smooth_loss(channel, img).
    flow (b,2,h,w) planar (FlowNet output), img (b,h,w,3) logical NHWC -> scalar.
    The reference's own body hands ``flow[:, i].unsqueeze(1)`` -- a (b,1,h,w) tensor -- to a function written for (b,h,w,1)
    and raises on the shape mismatch of its first multiply (it is only reachable from the commented-out flow stage,
    models/baseline.py:990-1059).  The INTENDED computation, defined and tested here: each flow channel is one more
    "depth-like" map, i.e. channel i goes in as (b,h,w,1); its gradient flows back into that channel, none to the image."""
    if flow.dim() != 4 or flow.shape[1] != 2:
        raise ValueError(f"flow_smooth_loss: flow must be (b,2,h,w), got {tuple(flow.shape)}")
    s = 0
    for i in range(2):
        s = s + smooth_loss(flow[:, i, :, :].unsqueeze(3), img)
    return s / 2


# ---- north-star compositions ----------------------------------------------------------------------
def inverse_warp(img, depth, pose, intrinsics, reverse_pose=False, padding_mode="zeros"):
    """img (N,h,w,3), depth (N,h,w), pose (N,6), intrinsics (N,3,3) ->
    (warped (N,h,w,3), valid (N,h,w)); == flow_warp(img, compute_rigid_flow(pose, depth, K, reverse_pose)).
    'zeros' = the reference sampler's semantics; 'border' is an extension (unpinned by the reference)."""
    flow = compute_rigid_flow(pose, depth, intrinsics, reverse_pose)
    return _Sampler.apply(img.float(), flow, True, padding_mode, True)      # (warped, valid) from one launch


_PLAN_CACHE: "dict[tuple, ops.FusedLossPlan]" = {}
_PYR_CACHE: "dict[tuple, tuple]" = {}
_PLAN_CACHE_MAX = 16


def _cached_plan(cfg, poses, K_ms, depth_pyr, tgt_pyr, src_pyr, need, want_K, want_img=False):
    """FusedLossPlan for this call SHAPE: shapes, strides, device, stream and options -- never buffer addresses, so a trainer
    whose depth / poses are fresh network outputs every step hits the cache.  The cached plan holds no caller tensor; its
    pointer fields are patched per call (ops.FusedLossPlan.rebind)."""
    def sig(t):
        return (tuple(t.shape), tuple(t.stride()))
    masks = cfg["masks"]
    key = (str(poses.device), torch.cuda.current_stream().cuda_stream, sig(poses), tuple(sig(t) for t in depth_pyr),
           tuple(sig(t) for t in tgt_pyr), tuple(sig(t) for t in src_pyr),
           cfg["alpha"], tuple(cfg["photo_weights"]), tuple(cfg["smooth_weights"]), need, want_K, want_img, cfg["padding_mode"],
           None if masks is None else tuple((sig(a), sig(b)) for a, b in masks), cfg["no_photo"], cfg["no_smooth"], cfg["disparity"])
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        plan = ops.FusedLossPlan(tgt_pyr, src_pyr, depth_pyr, poses, K_ms, cfg["alpha"], cfg["photo_weights"], cfg["smooth_weights"],
                                 want_grad=need, want_grad_K=want_K, padding_mode=cfg["padding_mode"], masks=masks,
                                 no_photo=cfg["no_photo"], no_smooth=cfg["no_smooth"], disparity=cfg["disparity"], keep_inputs=False,
                                 want_image_grads=want_img)
        _PLAN_CACHE[key] = plan
    plan.rebind(tgt_pyr, src_pyr, depth_pyr, poses, K_ms, masks)
    return plan


class _FusedLoss(torch.autograd.Function):
    """loss_total (differentiable w.r.t. poses, depth_pyr, K_ms) + the detail vector for logging.
    The gradients are produced in the same single pass as the loss, straight into tensors this call owns (no copies); the
    backward scales them by the upstream gradient in place -- or hands them out untouched when the caller declared the
    loss to be the root of the graph (cfg['unit_upstream'], as in ``loss_total.backward()``, models/baseline.py:1205)."""

    @staticmethod
    def forward(ctx, cfg, poses, K_ms, *rest):
        ns = cfg["ns"]
        depth_pyr = [t.detach() for t in rest[:ns]]
        tgt_pyr, src_pyr = rest[ns:2 * ns], rest[2 * ns:3 * ns]
        need = any(ctx.needs_input_grad)
        want_K = ctx.needs_input_grad[2]
        want_img = any(ctx.needs_input_grad[3 + ns:])          # images that require grad (never in the reference's trainer)
        tgt_pyr, src_pyr = [t.detach() for t in tgt_pyr], [t.detach() for t in src_pyr]
        poses_c = poses.detach()
        if not poses_c.is_contiguous():
            poses_c = poses_c.contiguous()
        plan = _cached_plan(cfg, poses_c, K_ms.detach(), depth_pyr, tgt_pyr, src_pyr, need, want_K, want_img)
        out, gposes, gK, gdepth = plan.fresh_outputs()
        gimg = (plan.grad_tgt_img + plan.grad_src_img) if want_img else []
        plan.run()
        if need:
            ctx.ns, ctx.has_K, ctx.unit, ctx.has_img = ns, gK is not None, bool(cfg["unit_upstream"]), want_img
            ctx.save_for_backward(gposes, *([gK] if gK is not None else []), *gdepth, *gimg)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g, _g_all):
        saved = ctx.saved_tensors
        ns = ctx.ns
        gposes = saved[0]
        gK = saved[1] if ctx.has_K else None
        o = 1 + (1 if ctx.has_K else 0)
        gdepth = list(saved[o:o + ns])
        gimg = list(saved[o + ns:]) if ctx.has_img else [None] * (2 * ns)
        if not ctx.unit:
            # one multi-tensor launch over tensors only this graph node refers to (in place: no copies)
            torch._foreach_mul_(gdepth + [gposes] + ([gK] if gK is not None else []) + (gimg if ctx.has_img else []), g)
        return (None, gposes if ctx.needs_input_grad[1] else None, gK, *gdepth, *gimg)


def geonet_loss(tgt_pyr: Sequence[torch.Tensor], src_pyr: Sequence[torch.Tensor], depth_pyr: Sequence[torch.Tensor],
                poses: torch.Tensor, K_ms: torch.Tensor, alpha: float = 0.85, loss_weight_rigid_warp: float = 1.0,
                loss_weight_disparity_smooth: float = 0.5, padding_mode: str = "zeros", masks=None,
                no_photo: bool = False, no_smooth: bool = False, disparity: bool = False, unit_upstream: bool = False):
    """Everything GeoNetModel.build_rigid_warp_flow + build_losses compute (models/baseline.py:867-970,
    1115-1161) in one fused pass.  tgt_pyr[s] (B,3,h,w), src_pyr[s] (S*B,3,h,w) source-major (any strides, e.g.
    the reference's NHWC views permuted back), depth_pyr[s] (B*(1+S),h,w), poses (B,S,6), K_ms (B,scales,3,3).
    disparity=True: depth_pyr holds the networks' DISPARITIES; the kernel forms depth = 1/disp itself (baseline.py:835) and
    the gradients come back with respect to the disparities (section 8(f2): saves the caller two element-wise passes).
    unit_upstream=True: the caller promises that loss_total is the root of the backward pass (upstream gradient exactly 1,
    as in the reference's ``loss_total.backward()``): the backward then returns the kernel's gradients without touching them.
    Returns (loss_total, detail) with detail = [total, rigid_warp, disp_smooth, then per scale mean fwd err,
    mean bwd err, smooth_loss]."""
    ns = len(tgt_pyr)
    S = poses.shape[1]
    cfg = dict(ns=ns, alpha=float(alpha), padding_mode=padding_mode, masks=masks, no_photo=no_photo, no_smooth=no_smooth,
               disparity=bool(disparity), unit_upstream=bool(unit_upstream),
               photo_weights=[loss_weight_rigid_warp * S / 2] * ns,
               smooth_weights=[loss_weight_disparity_smooth / 2 ** s for s in range(ns)])
    return _FusedLoss.apply(cfg, poses, K_ms, *depth_pyr, *tgt_pyr, *src_pyr)


def photometric_reconstruction_loss(tgt_pyr, src_pyr, intrinsics_ms, depth_pyr, poses, explainability_mask=None,
                                    alpha: float = 0.85, weight: float = 1.0, padding_mode: str = "zeros"):
    """North-star name: the reference's loss_rigid_warp (both directions, all scales) as one scalar.
    explainability_mask: optional list per scale of (mask_fwd, mask_bwd), each (S*B,h,w), multiplying the
    per-pixel error before the mean (extension: the reference's PoseNet has no mask head)."""
    total, _ = geonet_loss(tgt_pyr, src_pyr, depth_pyr, poses, intrinsics_ms, alpha, weight, 0.0, padding_mode,
                           explainability_mask, no_smooth=True)
    return total


def _frame_pyramids(tgt_u8, src_u8, ns: int):
    """uint8 frames as the DataLoader delivers them -- tgt_u8 (B,3,H,W), src_u8 (B,3*S,H,W) -- normalised to [-1,1] and turned
    into planar pyramids on the device (models/baseline.py:761-784 + utils_edited.py:89-112): tgt_pyr[s] (B,3,h,w), src_pyr[s]
    (S*B,3,h,w) source-major.  The buffers are kept per (device, stream, shape): they are consumed by kernels launched on
    the same stream (images take no gradient, nothing is saved for a backward), so the next call may overwrite them."""
    B = tgt_u8.shape[0]
    S = src_u8.shape[1] // 3
    H, W = tgt_u8.shape[2], tgt_u8.shape[3]
    key = ("pyr", str(tgt_u8.device), torch.cuda.current_stream().cuda_stream, tuple(tgt_u8.shape), tuple(src_u8.shape), ns)
    bufs = _PYR_CACHE.get(key)
    if bufs is None:
        if len(_PYR_CACHE) >= _PLAN_CACHE_MAX:
            _PYR_CACHE.pop(next(iter(_PYR_CACHE)))
        dev = tgt_u8.device
        bufs = ([torch.empty(B, 3, H >> s, W >> s, device=dev) for s in range(ns)],
                [torch.empty(S * B, 3, H >> s, W >> s, device=dev) for s in range(ns)])
        _PYR_CACHE[key] = bufs
    tgt_pyr, src_pyr = bufs
    ops.frames_to_pyramid(tgt_u8, ns, out=tgt_pyr)
    ops.frames_to_pyramid(src_u8, ns, out=src_pyr)      # (B,3S,H,W) -> source-major (S*B,3,h,w) levels (baseline.py:772-775), one launch
    return tgt_pyr, src_pyr


def _intrinsics_pyramid(intrinsics, ns: int):
    """(B,3,3) -> (B,scales,3,3) view of a scale-major buffer: K_s = K * M_s + E (fx, fy, cx, cy / 2^s, zeros and a one
    elsewhere: utils_edited.py:149-169); every K_ms[:, s] is a contiguous (B,3,3) block the C call can point at."""
    B = intrinsics.shape[0]
    key = ("K", str(intrinsics.device), torch.cuda.current_stream().cuda_stream, B, ns)
    bufs = _PYR_CACHE.get(key)
    if bufs is None:
        dev = intrinsics.device
        M = torch.zeros(ns, 1, 3, 3, device=dev)
        for s in range(ns):
            M[s, 0, 0, 0] = M[s, 0, 1, 1] = M[s, 0, 0, 2] = M[s, 0, 1, 2] = 1.0 / 2 ** s
        E = torch.zeros(1, 1, 3, 3, device=dev)
        E[0, 0, 2, 2] = 1.0
        bufs = (torch.empty(ns, B, 3, 3, device=dev), M, E)
        _PYR_CACHE[key] = bufs
    K_sb, M, E = bufs
    torch.mul(intrinsics.float().unsqueeze(0), M, out=K_sb)
    K_sb += E
    return K_sb.permute(1, 0, 2, 3)


def geonet_loss_from_frames(tgt_u8, src_u8, depth_pyr, poses, intrinsics, alpha: float = 0.85, loss_weight_rigid_warp: float = 1.0,
                            loss_weight_disparity_smooth: float = 0.5, disparity: bool = False, unit_upstream: bool = False):
    """The step as the reference's trainer sees it (models/baseline.py:746-792 + 867-986 + 1115-1161): uint8 frames as the
    DataLoader delivers them -- tgt_u8 (B,3,H,W), src_u8 (B,3*S,H,W) -- are normalised and turned into pyramids on the
    device (one kernel each), intrinsics (B,3,3) into their per-scale copies, then the fused loss runs.
    depth_pyr[s] (B*(1+S),h_s,w_s), poses (B,S,6).  Returns (loss_total, detail) like geonet_loss."""
    ns = len(depth_pyr)
    tgt_pyr, src_pyr = _frame_pyramids(tgt_u8, src_u8, ns)
    K_ms = _intrinsics_pyramid(intrinsics, ns)
    return geonet_loss(tgt_pyr, src_pyr, depth_pyr, poses, K_ms, alpha, loss_weight_rigid_warp, loss_weight_disparity_smooth,
                       disparity=disparity, unit_upstream=unit_upstream)
