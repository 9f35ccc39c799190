#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE (unmodified, imported from
/root/reference on CPU).  Run in the authoring container only:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these files
are what pins the oracle (tests/test_oracle_vs_golden.py) and, through it and directly,
the CUDA path (tests/test_gpu_*.py).  /root/reference does not exist on the GPU box;
nothing but this script reads it.

Reference entry points exercised (all fp32, CPU):
  utils/utils_edited.py : meshgrid, pose_vec2mat, compute_rigid_flow, flow_warp, bilinear_sampler, DSSIM (kats, cases);
                          scale_pyramid, compute_multi_scale_intrinsics, euler2mat, pixel2cam, cam2pixel (glue_kats)
  models/loss_functions.py : image_similarity, smooth_loss
composed exactly as models/baseline.py:895-970 (build_rigid_warp_flow) and :1125-1161
(build_losses) compose them.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    def _stub(name, **kw):
        m = types.ModuleType(name); m.__dict__.update(kw); sys.modules[name] = m; return m
    mpl = _stub("matplotlib", use=lambda *a, **k: None); mpl.pyplot = _stub("matplotlib.pyplot")
    sys.path.insert(0, "/root/reference")
    import utils.utils_edited as U
    import models.loss_functions as L
    return U, L


def reference_loss_path(U, L, inp, alpha=0.85, w_rigid=1.0, w_smooth=0.5, keep=False):
    """build_rigid_warp_flow + build_losses with the reference's own functions."""
    B, S, ns = inp.B, inp.S, inp.num_scales
    tgt_pyr = [t.permute(0, 2, 3, 1) for t in inp.tgt_pyr]           # NHWC views, as scale_pyramid returns
    src_pyr = [t.permute(0, 2, 3, 1) for t in inp.src_pyr]
    tgt_tile = [t.repeat(S, 1, 1, 1) for t in tgt_pyr]
    loss_rigid, loss_smooth = 0, 0
    rec = dict(fwd_flow=[], bwd_flow=[], fwd_warp=[], bwd_warp=[], fwd_err=[], bwd_err=[], fwd_mean=[], bwd_mean=[], smooth=[])
    for s in range(ns):
        fl_f, fl_b = [], []
        for k in range(S):
            fl_f.append(U.compute_rigid_flow(inp.poses[:, k, :], inp.depth_pyr[s][:B], inp.K_ms[:, s], False))
            fl_b.append(U.compute_rigid_flow(inp.poses[:, k, :], inp.depth_pyr[s][B * (k + 1):B * (k + 2)], inp.K_ms[:, s], True))
        fwd_flow, bwd_flow = torch.cat(fl_f, 0), torch.cat(fl_b, 0)
        fwd_warp = U.flow_warp(src_pyr[s], fwd_flow)
        bwd_warp = U.flow_warp(tgt_tile[s], bwd_flow)
        fwd_err = L.image_similarity(alpha, tgt_tile[s], fwd_warp)
        bwd_err = L.image_similarity(alpha, src_pyr[s], bwd_warp)
        fm, bm = torch.mean(fwd_err), torch.mean(bwd_err)
        sm = L.smooth_loss(inp.depth_pyr[s].unsqueeze(3), torch.cat((tgt_pyr[s], src_pyr[s]), 0))
        loss_rigid = loss_rigid + w_rigid * S / 2 * (fm + bm)
        loss_smooth = loss_smooth + w_smooth / (2 ** s) * sm
        rec["fwd_mean"].append(fm.item()); rec["bwd_mean"].append(bm.item()); rec["smooth"].append(sm.item())
        if keep:
            for k_, v in (("fwd_flow", fwd_flow), ("bwd_flow", bwd_flow), ("fwd_warp", fwd_warp), ("bwd_warp", bwd_warp),
                          ("fwd_err", fwd_err), ("bwd_err", bwd_err)):
                rec[k_].append(v.detach().numpy().copy())
    return loss_rigid, loss_smooth, rec


def run_case(U, L, name, keep, **gen):
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(**gen)
    for d in inp.depth_pyr:
        d.requires_grad_(True)
    inp.poses.requires_grad_(True)
    lr, ls, rec = reference_loss_path(U, L, inp, keep=keep)
    total = lr + ls
    total.backward()
    out = dict(loss_rigid_warp=np.float64(lr.item()), loss_disp_smooth=np.float64(ls.item()), loss_total=np.float64(total.item()),
               fwd_mean=np.array(rec["fwd_mean"]), bwd_mean=np.array(rec["bwd_mean"]), smooth=np.array(rec["smooth"]),
               grad_poses=inp.poses.grad.numpy().copy(),
               gen=np.array(repr(sorted(gen.items()))))
    for s, d in enumerate(inp.depth_pyr):
        g = d.grad.numpy()
        if keep:
            out[f"grad_depth_{s}"] = g.copy()
        out[f"grad_depth_{s}_l2"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out[f"grad_depth_{s}_sum"] = np.float64(g.astype(np.float64).sum())
        # a fixed pseudo-random projection: a cheap checksum sensitive to position
        rng = np.random.default_rng(1234 + s)
        out[f"grad_depth_{s}_proj"] = np.float64((g.astype(np.float64) * rng.standard_normal(g.shape)).sum())
    if keep:
        for k_ in ("fwd_flow", "bwd_flow", "fwd_warp", "bwd_warp", "fwd_err", "bwd_err"):
            for s, v in enumerate(rec[k_]):
                out[f"{k_}_{s}"] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss_total", out["loss_total"], "rigid", out["loss_rigid_warp"], "smooth", out["loss_disp_smooth"])


def kats(U, L):
    """Known-answer tables for the pieces (sampler edge semantics, grid tables, pose matrices)."""
    out = {}
    for n in (2, 3, 13, 16, 52, 104, 128, 208, 416, 480, 640, 1248, 2048):
        out[f"grid_{n}"] = ((torch.linspace(-1, 1, n) + 1) * 0.5 * (n - 1)).numpy()
    img = torch.arange(20.).view(1, 4, 5, 1)
    xs = [-2.3, -1.0, -0.5, 0.0, 0.25, 3.999, 4.0, 4.5, 6.0]
    ys = [-1.5, -0.25, 0.0, 1.0, 2.5, 2.999, 3.0, 3.7]
    tab = np.zeros((len(ys), len(xs)), np.float32)
    for i, y in enumerate(ys):
        for j, x in enumerate(xs):
            tab[i, j] = U.bilinear_sampler(img, torch.tensor([[[[x, y]]]])).item()
    out["sampler_xs"], out["sampler_ys"], out["sampler_tab"] = np.array(xs, np.float32), np.array(ys, np.float32), tab
    g = torch.Generator().manual_seed(7)
    vec = torch.randn(16, 6, generator=g) * torch.tensor([0.3, 0.3, 0.3, 0.5, 0.5, 0.5])
    out["pose_vec"] = vec.numpy()
    out["pose_mat"] = U.pose_vec2mat(vec).numpy()
    out["pose_mat_inv"] = torch.inverse(U.pose_vec2mat(vec)).numpy()
    # random-coordinate sampler case incl. far out-of-range coordinates, with gradients
    imgs = torch.rand(2, 9, 11, 3, generator=g)
    coords = (torch.rand(2, 7, 8, 2, generator=g) * torch.tensor([14.0, 12.0]) - 2.0).requires_grad_(True)
    imgs.requires_grad_(True)
    o = U.bilinear_sampler(imgs, coords)
    go = torch.randn(o.shape, generator=g)
    o.backward(go)
    out.update(bs_img=imgs.detach().numpy(), bs_coords=coords.detach().numpy(), bs_out=o.detach().numpy(), bs_gout=go.numpy(),
               bs_gcoords=coords.grad.numpy(), bs_gimg=imgs.grad.numpy())
    # image_similarity with arbitrary upstream gradient, both inputs differentiable
    x = (torch.rand(2, 10, 12, 3, generator=g) * 2 - 1).requires_grad_(True)
    y = (torch.rand(2, 10, 12, 3, generator=g) * 2 - 1).requires_grad_(True)
    e = L.image_similarity(0.85, x, y)
    ge = torch.randn(e.shape, generator=g)
    e.backward(ge)
    out.update(is_x=x.detach().numpy(), is_y=y.detach().numpy(), is_err=e.detach().numpy(), is_gerr=ge.numpy(),
               is_gx=x.grad.numpy(), is_gy=y.grad.numpy(), is_dssim=U.DSSIM(x.detach(), y.detach()).numpy())
    # smooth_loss
    d = (torch.rand(3, 10, 12, 1, generator=g) + 0.1).requires_grad_(True)
    im = torch.rand(3, 10, 12, 3, generator=g) * 2 - 1
    sl = L.smooth_loss(d, im)
    sl.backward()
    out.update(sm_depth=d.detach().numpy(), sm_img=im.numpy(), sm_loss=np.float64(sl.item()), sm_gdepth=d.grad.numpy())
    # compute_rigid_flow incl. the matrices it builds internally
    Bk, hk, wk = 3, 12, 20
    Kk = torch.tensor([[0.58 * wk, 0, 0.5 * wk], [0, 1.92 * hk, 0.5 * hk], [0, 0, 1]]).repeat(Bk, 1, 1)
    Kk[1, 0, 1] = 0.3  # skew on one sample: exercises the general 3x3 inverse
    pk = (torch.randn(Bk, 6, generator=g) * torch.tensor([0.1, 0.1, 0.1, 0.05, 0.05, 0.05])).requires_grad_(True)
    dk = (1.0 / (torch.rand(Bk, hk, wk, generator=g) * 9 + 0.5)).requires_grad_(True)
    for rev in (False, True):
        fl = U.compute_rigid_flow(pk, dk, Kk, rev)
        gfl = torch.randn(fl.shape, generator=g)
        gp, gd = torch.autograd.grad(fl, (pk, dk), gfl)
        tag = "rev" if rev else "fwd"
        out.update({f"rf_{tag}_flow": fl.detach().numpy(), f"rf_{tag}_gflow": gfl.numpy(), f"rf_{tag}_gpose": gp.numpy(),
                    f"rf_{tag}_gdepth": gd.numpy()})
    out.update(rf_K=Kk.numpy(), rf_pose=pk.detach().numpy(), rf_depth=dk.detach().numpy(), rf_Kinv=torch.inverse(Kk).numpy())
    np.savez_compressed(os.path.join(HERE, "kats.npz"), **out)
    print("kats written:", len(out), "arrays")


def glue_kats(U, L):
    """Round 2: known answers for the caller-side glue the drop-in modules restate in torch:
    scale_pyramid, compute_multi_scale_intrinsics, euler2mat, pose_vec2mat, meshgrid, pixel2cam, cam2pixel -- each one
    really called on the imported reference."""
    out = {}
    g = torch.Generator().manual_seed(21)
    img = torch.rand(3, 3, 48, 80, generator=g) * 2 - 1
    out["sp_img"] = img.numpy()
    for s, lvl in enumerate(U.scale_pyramid(img, 4)):                    # NHWC views of NCHW memory
        out[f"sp_level_{s}"] = np.ascontiguousarray(lvl.numpy())
    K = torch.tensor([[241.67, 0.0, 204.17], [0.0, 246.28, 59.0], [0.0, 0.0, 1.0]]).repeat(3, 1, 1)
    K = K * (1 + 0.1 * torch.rand(3, 1, 1, generator=g))
    K[:, 2, 2] = 1.0
    out["msi_K"] = K.numpy()
    out["msi_out"] = U.compute_multi_scale_intrinsics(K, 4).numpy()
    ang = torch.randn(5, 3, generator=g) * 0.7
    out["e2m_zyx"] = ang.numpy()
    out["e2m_out"] = U.euler2mat(ang[:, 0], ang[:, 1], ang[:, 2]).squeeze(1).numpy()
    for b, h, w, hom in ((2, 6, 9, True), (1, 5, 7, False)):
        out[f"mg_{h}x{w}_{int(hom)}"] = np.ascontiguousarray(U.meshgrid(b, h, w, hom).numpy())
    b, h, w = 2, 6, 9
    depth = 1.0 / (torch.rand(b, h, w, generator=g) * 9 + 0.5)
    Kc = torch.tensor([[0.58 * w, 0.1, 0.5 * w], [0, 1.92 * h, 0.5 * h], [0, 0, 1]]).repeat(b, 1, 1)
    pix = U.meshgrid(b, h, w)
    cam = U.pixel2cam(depth, pix, Kc)
    out.update(p2c_depth=depth.numpy(), p2c_K=Kc.numpy(), p2c_out=cam.numpy())
    cam3 = U.pixel2cam(depth, pix, Kc, False)
    out["p2c_out_nohom"] = cam3.numpy()
    vec = torch.randn(b, 6, generator=g) * torch.tensor([0.1, 0.1, 0.1, 0.05, 0.05, 0.05])
    K4 = torch.zeros(b, 4, 4); K4[:, :3, :3] = Kc; K4[:, 3, 3] = 1.0
    proj = torch.matmul(K4, U.pose_vec2mat(vec))
    out.update(c2p_vec=vec.numpy(), c2p_proj=proj.numpy(), c2p_out=np.ascontiguousarray(U.cam2pixel(cam, proj).numpy()))
    # flow_smooth_loss with its INTENDED semantics (the reference's body raises: it feeds (b,1,h,w) slices to a (b,h,w,1)
    # function): every flow channel as a (b,h,w,1) map through the reference's own smooth_loss, averaged
    g2 = torch.Generator().manual_seed(33)
    flow = (torch.randn(2, 2, 10, 14, generator=g2) * 3).requires_grad_(True)
    fimg = torch.rand(2, 10, 14, 3, generator=g2) * 2 - 1
    fs = sum(L.smooth_loss(flow[:, i, :, :].unsqueeze(3), fimg) for i in range(2)) / 2
    fs.backward()
    out.update(fsl_flow=flow.detach().numpy(), fsl_img=fimg.numpy(), fsl_loss=np.float64(fs.item()), fsl_gflow=flow.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "glue_kats.npz"), **out)
    print("glue_kats written:", len(out), "arrays")


def full_size_cases(U, L):
    """Round 2: the BASELINE shapes at their FULL bench sizes through the reference itself (scalars, pose gradients and
    position-sensitive checksums of the depth gradients): config 2 (B=32), config 3 shape (NYU, B=8) and config 4's per-GPU
    shard (376x1248, S=4, B=8: N*h*w = 15.0 M stays below the reference's 2^24 float-index limit)."""
    run_case(U, L, "config2_smooth", False, B=32, S=2, H=128, W=416, num_scales=4, seed=12, smooth=True, motion="small")
    run_case(U, L, "nyu_b8_smooth", False, B=8, S=2, H=480, W=640, num_scales=4, seed=13, smooth=True, motion="small", intrinsics="nyu")
    run_case(U, L, "hires_b8_smooth", False, B=8, S=4, H=376, W=1248, num_scales=4, seed=14, smooth=True, motion="small")


def network_kats(U, L):
    """Round 2: the reference's DispNetS / PoseNet (models/DispNetS.py, models/PoseNet.py) with seeded xavier initialisation on
    small seeded inputs -- pins the architecture, the head scaling and the initialisation order of snde_b200/networks.py."""
    from models.DispNetS import DispNetS
    from models.PoseNet import PoseNet
    out = {}
    torch.manual_seed(0); d = DispNetS(); d.init_weights(); d.train()
    torch.manual_seed(1); p = PoseNet(2); p.init_weight(); p.train()
    out["disp_params"] = np.int64(sum(t.numel() for t in d.parameters()))
    out["pose_params"] = np.int64(sum(t.numel() for t in p.parameters()))
    out["disp_keys"] = np.array(list(d.state_dict().keys()))
    out["pose_keys"] = np.array(list(p.state_dict().keys()))
    out["disp_abs_sums"] = np.array([float(t.double().abs().sum()) for t in d.state_dict().values()])
    out["pose_abs_sums"] = np.array([float(t.double().abs().sum()) for t in p.state_dict().values()])
    g = torch.Generator().manual_seed(5)
    for tag, shape in (("a", (2, 3, 128, 160)), ("b", (1, 3, 72, 104))):          # b: not divisible by 128 (decoder crops)
        x = torch.rand(shape, generator=g) * 2 - 1
        out[f"disp_in_{tag}"] = x.numpy()
        with torch.no_grad():
            for i, t in enumerate(d(x)):
                out[f"disp_out_{tag}_{i}"] = t.numpy()
    xp = torch.rand(2, 9, 128, 160, generator=g) * 2 - 1
    out["pose_in"] = xp.numpy()
    with torch.no_grad():
        out["pose_out"] = p(xp).numpy()
    np.savez_compressed(os.path.join(HERE, "network_kats.npz"), **out)
    print("network_kats written:", len(out), "arrays")


def train_step_kats(U, L):
    """Round 2: ONE iteration of the reference's trainer (GeoNetModel: iter_data_preparation, build_dispnet, build_posenet,
    build_rigid_warp_flow, build_losses, backward, Adam step -- models/baseline.py:1170-1206) on seeded uint8 frames:
    losses, network outputs and per-parameter gradient norms.  Pins snde_b200/trainer.py (step-level parity)."""
    import argparse
    import tempfile

    class _SW:
        def __init__(self, *a, **k): pass
        def add_scalar(self, *a, **k): pass
        def add_images(self, *a, **k): pass
    for name, kw in (("tensorboardX", dict(SummaryWriter=_SW)), ("imageio", dict(imread=lambda *a, **k: None))):
        m = types.ModuleType(name); m.__dict__.update(kw); sys.modules[name] = m
    import models.baseline as MB
    MB.n_iter = 1
    B, S, H, W, ns, seed = 2, 2, 128, 160, 4, 8964
    tmp = tempfile.mkdtemp()
    args = argparse.Namespace(is_train=1, train_flow=False, batch_size=B, num_source=S, num_scales=ns, sequence_length=S + 1,
                              img_height=H, img_width=W, simi_alpha=0.85, loss_weight_rigid_warp=1.0, loss_weight_disparity_smooth=0.5,
                              graphs_dir=tmp, ckpt_dir=tmp, learning_rate=2e-4, momentum=0.9, beta=0.999, weight_decay=0.0)
    torch.manual_seed(seed)
    model = MB.GeoNetModel(args, torch.device("cpu"))
    model.disp_net.train(); model.pose_net.train()
    g = torch.Generator().manual_seed(77)
    # band-limited frames (uint8), a KITTI-like intrinsics matrix
    low = torch.rand(B, 3 * (1 + S), H // 8, W // 8, generator=g)
    frames = (torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1) * 255).round().to(torch.uint8)
    tgt_u8, src_u8 = frames[:, :3].contiguous(), frames[:, 3:].contiguous()
    K = torch.tensor([[0.58 * W, 0, 0.5 * W], [0, 1.92 * H, 0.5 * H], [0, 0, 1]]).repeat(B, 1, 1)
    opt = torch.optim.Adam([{"params": v.parameters(), "lr": args.learning_rate} for v in model.nets.values()],
                           betas=(args.momentum, args.beta), weight_decay=args.weight_decay)
    model.iter_data_preparation((tgt_u8, src_u8, K))
    model.build_dispnet(); model.build_posenet(); model.build_rigid_warp_flow(); model.build_losses()
    opt.zero_grad(); model.loss_total.backward()
    out = dict(B=np.int64(B), S=np.int64(S), H=np.int64(H), W=np.int64(W), seed=np.int64(seed), tgt_u8=tgt_u8.numpy(), src_u8=src_u8.numpy(), K=K.numpy(),
               loss_total=np.float64(model.loss_total.item()), loss_rigid_warp=np.float64(model.loss_rigid_warp.item()),
               loss_disp_smooth=np.float64(model.loss_disp_smooth.item()), poses=model.poses.detach().numpy(),
               disp0_mean=np.float64(model.disparities[0].double().mean().item()),
               disp_grad_l2=np.array([float(p.grad.double().norm()) for p in model.disp_net.parameters()]),
               pose_grad_l2=np.array([float(p.grad.double().norm()) for p in model.pose_net.parameters()]))
    opt.step()
    out["disp_abs_sum_after_step"] = np.array([float(p.detach().double().abs().sum()) for p in model.disp_net.parameters()])
    out["pose_abs_sum_after_step"] = np.array([float(p.detach().double().abs().sum()) for p in model.pose_net.parameters()])
    np.savez_compressed(os.path.join(HERE, "train_step_kats.npz"), **out)
    print("train_step_kats written: loss_total", out["loss_total"], "rigid", out["loss_rigid_warp"], "smooth", out["loss_disp_smooth"])


def image_grad_kats(U, L):
    """Round 2 (section 8(f3)): what the reference's autograd gives IMAGES that require grad -- d loss_rigid_warp / d tgt_pyr[s],
    d src_pyr[s] through image_similarity (dense) and bilinear_sampler's index_select (scatter) -- on the small cases."""
    from snde_b200.synthetic import make_inputs
    out = {}
    for name, gen in (("small_smooth", dict(B=2, S=2, H=32, W=64, num_scales=3, seed=1, smooth=True, motion="small")),
                      ("small_s1_nyu", dict(B=1, S=1, H=30, W=44, num_scales=2, seed=3, smooth=True, motion="large", intrinsics="nyu"))):
        inp = make_inputs(**gen)
        for t in inp.tgt_pyr + inp.src_pyr:
            t.requires_grad_(True)
        lr, ls, _ = reference_loss_path(U, L, inp)
        lr.backward()
        out[f"{name}_loss_rigid_warp"] = np.float64(lr.item())
        for s in range(inp.num_scales):
            out[f"{name}_grad_tgt_{s}"] = inp.tgt_pyr[s].grad.numpy().copy()       # planar (B,3,h,w)
            out[f"{name}_grad_src_{s}"] = inp.src_pyr[s].grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "image_grad_kats.npz"), **out)
    print("image_grad_kats written:", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    U, L = import_reference()
    if "--networks" in sys.argv:
        network_kats(U, L)
        sys.exit(0)
    if "--image-grads" in sys.argv:
        image_grad_kats(U, L)
        sys.exit(0)
    if "--train-step" in sys.argv:
        train_step_kats(U, L)
        sys.exit(0)
    if "--round2" in sys.argv:          # only the fixtures added in round 2 (the round-1 files stay byte-identical)
        network_kats(U, L)
        train_step_kats(U, L)
        image_grad_kats(U, L)
        glue_kats(U, L)
        full_size_cases(U, L)
        sys.exit(0)
    kats(U, L)
    # small cases with every intermediate kept (a few hundred KB each)
    run_case(U, L, "small_smooth", True, B=2, S=2, H=32, W=64, num_scales=3, seed=1, smooth=True, motion="small")
    run_case(U, L, "small_noise_large_motion", True, B=2, S=2, H=24, W=40, num_scales=2, seed=2, smooth=False, motion="large")
    run_case(U, L, "small_s1_nyu", True, B=1, S=1, H=30, W=44, num_scales=2, seed=3, smooth=True, motion="large", intrinsics="nyu")
    # BASELINE config 1 (the reference's CPU-runnable case), scalars + checksums only
    run_case(U, L, "config1_noise", False, B=4, S=2, H=128, W=416, num_scales=4, seed=0, smooth=False, motion="small")
    run_case(U, L, "config1_smooth", False, B=4, S=2, H=128, W=416, num_scales=4, seed=0, smooth=True, motion="small")
    glue_kats(U, L)
    full_size_cases(U, L)
    network_kats(U, L)
    train_step_kats(U, L)
    image_grad_kats(U, L)
