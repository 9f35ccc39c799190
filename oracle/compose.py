"""Oracle composition of the loss path -- TEST INFRASTRUCTURE ONLY.

Restates what ``GeoNetModel.build_rigid_warp_flow`` and ``build_losses`` do with the
per-function oracle (reference ``models/baseline.py:895-970`` and ``:1125-1161``):

  for every scale s, source k:   fwd flow from TARGET depth with pose k,
                                  bwd flow from SOURCE-k depth with inverse pose k
  fwd_warp = flow_warp(src_pyr[s], fwd_flow)      bwd_warp = flow_warp(tgt_tile[s], bwd_flow)
  fwd_err  = image_similarity(a, tgt_tile, fwd_warp)   bwd_err = image_similarity(a, src_pyr, bwd_warp)
  loss_rigid_warp  = sum_s w_r * S/2 * (mean(fwd_err) + mean(bwd_err))
  loss_disp_smooth = sum_s w_s / 2^s * smooth_loss(depth_s, cat(tgt_s, src_s))

and the matching backward (what ``loss_total.backward()`` produces for depth, poses and,
as an extension the reference never exercises, the intrinsics).
"""
from __future__ import annotations

import numpy as np

from . import Oracle


def multi_scale_intrinsics(orc: Oracle, K, num_scales):
    """reference utils/utils_edited.py:149-169: fx,fy,cx,cy / 2^s, other entries 0/1."""
    K = np.asarray(K, orc.dt)
    B = K.shape[0]
    out = np.zeros((B, num_scales, 3, 3), orc.dt)
    for s in range(num_scales):
        out[:, s, 0, 0] = K[:, 0, 0] / orc.dt(2 ** s)
        out[:, s, 1, 1] = K[:, 1, 1] / orc.dt(2 ** s)
        out[:, s, 0, 2] = K[:, 0, 2] / orc.dt(2 ** s)
        out[:, s, 1, 2] = K[:, 1, 2] / orc.dt(2 ** s)
        out[:, s, 2, 2] = 1
    return out


def geonet_loss(orc: Oracle, tgt_pyr, src_pyr, depth_pyr, poses, K_ms, alpha=0.85, w_rigid=1.0,
                w_smooth=0.5, want_grad=True, keep_maps=False, masks=None, want_grad_images=False):
    """tgt_pyr[s] (B,h,w,3); src_pyr[s] (S*B,h,w,3) source-major; depth_pyr[s] (B(1+S),h,w);
    poses (B,S,6); K_ms (B,scales,3,3).  Returns a dict of losses (+ grads, + maps)."""
    num_scales = len(tgt_pyr)
    poses = np.asarray(poses, orc.dt)
    B, S = poses.shape[:2]
    N = B * S
    out = {"fwd_mean": [], "bwd_mean": [], "smooth": []}
    if keep_maps:
        out.update(fwd_warp=[], bwd_warp=[], fwd_err=[], bwd_err=[], fwd_flow=[], bwd_flow=[])
    gposes = np.zeros((B, S, 6), np.float64)
    gK_ms = np.zeros((B, num_scales, 3, 3), np.float64)
    gdepth_pyr = []
    loss_rigid, loss_smooth = 0.0, 0.0
    for s in range(num_scales):
        tgt = np.asarray(tgt_pyr[s], orc.dt)
        src = np.asarray(src_pyr[s], orc.dt)
        depth = np.asarray(depth_pyr[s], orc.dt)
        _, h, w, C = tgt.shape
        K = np.ascontiguousarray(K_ms[:, s])
        Pf, Pb, Kinv = orc.pose_prep(poses, K)
        Kinv_n = np.tile(Kinv, (S, 1, 1))                      # (N,3,3) source-major
        d_fwd = np.tile(depth[:B], (S, 1, 1))                  # target depth for every source
        d_bwd = depth[B:]                                      # source depths, already source-major
        tgt_tile = np.tile(tgt, (S, 1, 1, 1))
        fwd_flow = orc.rigid_flow(d_fwd, Kinv_n, Pf)
        bwd_flow = orc.rigid_flow(d_bwd, Kinv_n, Pb)
        fwd_coords, bwd_coords = orc.flow_to_coords(fwd_flow), orc.flow_to_coords(bwd_flow)
        fwd_warp = orc.bilinear_sampler(src, fwd_coords)
        bwd_warp = orc.bilinear_sampler(tgt_tile, bwd_coords)
        fwd_err = orc.image_similarity(alpha, tgt_tile, fwd_warp)
        bwd_err = orc.image_similarity(alpha, src, bwd_warp)
        if masks is not None:   # extension: explainability weights (S*B,h,w) multiply the error before the mean
            mf = np.asarray(masks[s][0], orc.dt)[..., None]; mb = np.asarray(masks[s][1], orc.dt)[..., None]
        else:
            mf = mb = None
        fm = orc.mean(fwd_err * mf) if mf is not None else orc.mean(fwd_err)
        bm = orc.mean(bwd_err * mb) if mb is not None else orc.mean(bwd_err)
        img_cat = np.concatenate([tgt, src], 0)
        sm = orc.smooth_loss(depth, img_cat)
        out["fwd_mean"].append(fm); out["bwd_mean"].append(bm); out["smooth"].append(sm)
        loss_rigid += w_rigid * S / 2 * (fm + bm)
        loss_smooth += w_smooth / (2 ** s) * sm
        if keep_maps:
            out["fwd_warp"].append(fwd_warp); out["bwd_warp"].append(bwd_warp)
            out["fwd_err"].append(fwd_err); out["bwd_err"].append(bwd_err)
            out["fwd_flow"].append(fwd_flow); out["bwd_flow"].append(bwd_flow)
        if want_grad:
            g = orc.dt(w_rigid * S / 2 / fwd_err.size)
            gerr = np.full(fwd_err.shape, g, orc.dt)
            gx_f, gy_f = orc.image_similarity_bwd(alpha, tgt_tile, fwd_warp, gerr * mf if mf is not None else gerr)
            gx_b, gy_b = orc.image_similarity_bwd(alpha, src, bwd_warp, gerr * mb if mb is not None else gerr)
            gc_f, gi_f = orc.bilinear_sampler_bwd(src, fwd_coords, gy_f, want_gimg=want_grad_images)
            gc_b, gi_b = orc.bilinear_sampler_bwd(tgt_tile, bwd_coords, gy_b, want_gimg=want_grad_images)
            if want_grad_images:
                # what autograd gives images that require grad (extension; the reference's images are data): the dense part through
                # image_similarity's x argument + the scatter part through index_select (utils_edited.py:473-476); target
                # copies (tgt_tile = tgt.repeat(S)) sum over the S sources.  smooth_loss gives the images nothing.
                out.setdefault("grad_tgt_img", []).append(
                    (gx_f.astype(np.float64) + gi_b.astype(np.float64)).reshape(S, B, h, w, C).sum(0).astype(orc.dt))
                out.setdefault("grad_src_img", []).append((gi_f.astype(np.float64) + gx_b.astype(np.float64)).astype(orc.dt))
            gd_f, gPf, gKi_f = orc.rigid_flow_bwd(d_fwd, Kinv_n, Pf, gc_f)
            gd_b, gPb, gKi_b = orc.rigid_flow_bwd(d_bwd, Kinv_n, Pb, gc_b)
            gdepth = np.zeros(depth.shape, np.float64)
            gdepth[:B] = gd_f.reshape(S, B, h, w).astype(np.float64).sum(0)
            gdepth[B:] = gd_b
            gdepth += orc.smooth_loss_bwd(depth, img_cat, w_smooth / (2 ** s))
            gdepth_pyr.append(gdepth.astype(orc.dt))
            gKinv = (gKi_f + gKi_b).reshape(S, B, 3, 3).sum(0)
            gp, gK = orc.pose_prep_bwd(poses, K, gPf, gPb, gKinv)
            gposes += gp
            gK_ms[:, s] = gK
    out["loss_rigid_warp"] = loss_rigid
    out["loss_disp_smooth"] = loss_smooth
    out["loss_total"] = loss_rigid + loss_smooth
    if want_grad:
        out["grad_depth"] = gdepth_pyr
        out["grad_poses"] = gposes
        out["grad_K_ms"] = gK_ms
    return out
