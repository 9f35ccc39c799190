#!/usr/bin/env python
"""What a BY-NAME drop-in (only the two import lines of models/baseline.py change) runs: the reference's
own call sequence -- compute_rigid_flow x 2*S*scales, flow_warp x 2*scales, image_similarity x 2*scales, smooth_loss x scales, then
loss_total.backward() (models/baseline.py:895-970, 1125-1161) -- on the reference-signature CUDA ops, at BASELINE config 2.
Also: the pyramid prep kernel alone (achieved GB/s against its algorithmic bytes) and the fused op through the same autograd API.
Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snde_b200 import ops
from snde_b200.loss_functions import geonet_loss, image_similarity, smooth_loss
from snde_b200.synthetic import make_inputs
from snde_b200.utils_edited import compute_rigid_flow, flow_warp

dev = torch.device("cuda:0")
cfg = dict(B=32, S=2, H=128, W=416, num_scales=4)
inp = make_inputs(seed=3, **cfg).to(dev)
B, S, ns = inp.B, inp.S, inp.num_scales
tgt_pyr = [t.permute(0, 2, 3, 1) for t in inp.tgt_pyr]          # NHWC views of NCHW memory, as scale_pyramid returns them
src_pyr = [t.permute(0, 2, 3, 1) for t in inp.src_pyr]


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def by_name():
    depth = [d.detach().requires_grad_(True) for d in inp.depth_pyr]
    poses = inp.poses.detach().requires_grad_(True)
    tgt_tile = [t.repeat(S, 1, 1, 1) for t in tgt_pyr]                                   # baseline.py:781-784
    rigid, smooth = 0, 0
    for s in range(ns):
        ff = torch.cat([compute_rigid_flow(poses[:, k, :], depth[s][:B], inp.K_ms[:, s], False) for k in range(S)], 0)
        fb = torch.cat([compute_rigid_flow(poses[:, k, :], depth[s][B * (k + 1):B * (k + 2)], inp.K_ms[:, s], True) for k in range(S)], 0)
        fw, bw = flow_warp(src_pyr[s], ff), flow_warp(tgt_tile[s], fb)
        fe, be = image_similarity(0.85, tgt_tile[s], fw), image_similarity(0.85, src_pyr[s], bw)
        rigid = rigid + 1.0 * S / 2 * (fe.mean() + be.mean())
        smooth = smooth + 0.5 / 2 ** s * smooth_loss(depth[s].unsqueeze(3), torch.cat((tgt_pyr[s], src_pyr[s]), 0))
    (rigid + smooth).backward()
    return rigid + smooth


def fused():
    depth = [d.detach().requires_grad_(True) for d in inp.depth_pyr]
    poses = inp.poses.detach().requires_grad_(True)
    total, _ = geonet_loss(inp.tgt_pyr, inp.src_pyr, depth, poses, inp.K_ms, unit_upstream=True)
    total.backward()
    return total


l_by, l_fused = float(by_name()), float(fused())
ms_by, ms_fused = timed(by_name), timed(fused, n=100)
# the pyramid prep alone: target block (B,3,H,W) + source block (B,3S,H,W), two launches
g = torch.Generator().manual_seed(1)
tgt_u8 = torch.randint(0, 256, (B, 3, cfg["H"], cfg["W"]), generator=g, dtype=torch.uint8).to(dev)
src_u8 = torch.randint(0, 256, (B, 3 * S, cfg["H"], cfg["W"]), generator=g, dtype=torch.uint8).to(dev)
t_out = [torch.empty(B, 3, cfg["H"] >> s, cfg["W"] >> s, device=dev) for s in range(ns)]
s_out = [torch.empty(S * B, 3, cfg["H"] >> s, cfg["W"] >> s, device=dev) for s in range(ns)]
def prep():
    ops.frames_to_pyramid(tgt_u8, ns, out=t_out); ops.frames_to_pyramid(src_u8, ns, out=s_out)
ms_prep = timed(prep, n=200)
px_ch = B * (1 + S) * 3 * cfg["H"] * cfg["W"]
prep_bytes = px_ch * (1 + 4 * sum(0.25 ** s for s in range(ns)))
warped = inp.warped_pixels()
print(json.dumps({
    "config": "config2: B=32, S=2, 128x416, 4 scales, one B200",
    "by_name_path": {"ms_per_fwd_bwd": round(ms_by, 3), "warped_px_per_s": warped / (ms_by * 1e-3), "loss": l_by,
                     "what": "reference call sequence on the reference-signature CUDA ops through autograd (about 150 op launches + torch glue)"},
    "fused_op_autograd": {"ms_per_fwd_bwd": round(ms_fused, 4), "warped_px_per_s": warped / (ms_fused * 1e-3), "loss": l_fused,
                          "what": "geonet_loss(...).backward() with fresh leaf tensors every call (plan cache hit, gradients owned by the call)"},
    "speedup_fused_over_by_name": round(ms_by / ms_fused, 1),
    "frames_to_pyramid": {"us_per_step": round(ms_prep * 1e3, 2), "launches": 2, "algorithmic_bytes": int(prep_bytes),
                          "achieved_gbs": round(prep_bytes / (ms_prep * 1e-3) / 1e9, 1),
                          "what": "uint8 (B,3,H,W) + (B,3S,H,W) -> both 4-level fp32 pyramids; 1 B read + 5.31 B written per pixel and channel"}}))
