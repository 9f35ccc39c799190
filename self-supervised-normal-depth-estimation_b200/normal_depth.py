"""Depth <-> normal consistency (BASELINE config 3): GeoNet's depth-to-normal / normal-to-depth kernel
regressions and a consistency loss on them, backed by libsnde.so (csrc/snde_dn.cu).  CUDA only.

The reference has these only as inference code inside ``NNET.forward`` (models/baseline.py:350-446, 495-547), with a scrambled
patch extraction, a hard-coded batch of 4 and no loss; what is implemented here is the intended computation (include/snde.h
states it formula by formula; oracle/normal_depth.py is its CPU restatement and the parity target).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from . import _abi
from .ops import _require_cuda, _stream, img_nchw, map_nhw

_grids: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}


def camera_grid(h: int, w: int, device, xmax: float = 0.6, ymax: float = 0.4):
    """the reference's normalised camera grid (models/baseline.py:308-317): linspace(-xmax, xmax, w), linspace(-ymax, ymax, h)"""
    key = (h, w, str(device), float(xmax), float(ymax))
    g = _grids.get(key)
    if g is None:
        g = (torch.linspace(-xmax, xmax, w).to(device), torch.linspace(-ymax, ymax, h).to(device))
        _grids[key] = g
    return g


def _run_fwd(depth, normal, k, rate, thresh, xmax, ymax, w_depth, w_normal, clamp_max, want_maps, want_sq):
    _require_cuda(depth, normal)
    B, h, w = depth.shape
    if normal.shape != (B, 3, h, w):
        raise ValueError("depth must be (B,h,w) and normal (B,3,h,w)")
    dev = depth.device
    xg, yg = camera_grid(h, w, dev, xmax, ymax)
    a = _abi.SndeDnArgs()
    a.B, a.h, a.w, a.k, a.rate = B, h, w, int(k), int(rate)
    a.thresh, a.clamp_max, a.w_depth, a.w_normal = float(thresh), float(clamp_max), float(w_depth), float(w_normal)
    a.depth, a.normal = map_nhw(depth), img_nchw(normal)
    a.xgrid, a.ygrid = xg.data_ptr(), yg.data_ptr()
    n_out = torch.empty(B, 3, h, w, device=dev) if want_maps else None
    d_out = torch.empty(B, h, w, device=dev) if want_maps else None
    if want_maps:
        a.normal_from_depth, a.depth_from_normal = img_nchw(n_out), map_nhw(d_out)
    loss = torch.empty(3, device=dev)
    sq = torch.empty(B, h, w, 2, device=dev) if want_sq else None
    sums = torch.empty(2, dtype=torch.float64, device=dev)
    a.loss_out, a.sums = loss.data_ptr(), sums.data_ptr()
    a.sq = sq.data_ptr() if want_sq else None
    _abi.check(_abi.lib().snde_dn_consistency_fwd(C.byref(a), _stream()))
    return a, loss, n_out, d_out, sq, (xg, yg, sums)


def depth_normal_maps(depth, normal, k=9, rate=4, thresh=0.95, xmax=0.6, ymax=0.4, clamp_max=10.0):
    """-> (normal_from_depth (B,3,h,w), depth_from_normal (B,h,w)): D2N and N2D of the reference's NNET.forward (intended form)"""
    _, _, n_out, d_out, _, _ = _run_fwd(depth.detach().float(), normal.detach().float(), k, rate, thresh, xmax, ymax, 1.0, 1.0, clamp_max, True, False)
    return n_out, d_out


class _Consistency(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, normal, cfg):
        d, n = depth.detach(), normal.detach()
        need = ctx.needs_input_grad[0]
        a, loss, _, _, sq, keep = _run_fwd(d, n, cfg["k"], cfg["rate"], cfg["thresh"], cfg["xmax"], cfg["ymax"], cfg["w_depth"], cfg["w_normal"],
                                           cfg["clamp_max"], False, need)
        if need:
            ctx.args, ctx.keep = a, keep
            ctx.save_for_backward(d, n, sq)
        ctx.mark_non_differentiable(loss)
        return loss[0], loss

    @staticmethod
    def backward(ctx, g, _gall):
        d, n, sq = ctx.saved_tensors
        gd = torch.empty(d.shape, device=d.device)
        gs = g.reshape(1).contiguous().float()
        _abi.check(_abi.lib().snde_dn_consistency_bwd(C.byref(ctx.args), gs.data_ptr(), map_nhw(gd), _stream()))
        return gd, None, None


def depth_normal_consistency_loss(depth, normal, k=9, rate=4, thresh=0.95, xmax=0.6, ymax=0.4, w_depth=1.0, w_normal=1.0, clamp_max=10.0):
    """depth (B,h,w), normal (B,3,h,w) -> (loss, detail = [total, mean|d - N2D|, mean(1 - n . D2N)]).
    Differentiable with respect to depth (normals are data; the normal term is reported, it carries no gradient)."""
    cfg = dict(k=k, rate=rate, thresh=thresh, xmax=xmax, ymax=ymax, w_depth=w_depth, w_normal=w_normal, clamp_max=clamp_max)
    return _Consistency.apply(depth.float(), normal.float(), cfg)
