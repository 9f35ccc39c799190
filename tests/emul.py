"""Host-emulation harness (tests only): runs the fused loss strip program of csrc/snde_strip.cuh on the
CPU through tests/host_emul/emul.cpp (32 fibers per warp), with numpy arrays behind the very same SndeLossArgs struct."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
DEFS = os.environ.get("SNDE_EMUL_DEFS", "").split()          # tuning experiments only (same switches as SNDE_NVCC_DEFS)
OUT = os.path.join(ROOT, "tests", "_build", "libsnde_emul%s.so" % ("_" + "".join(c for c in "".join(DEFS) if c.isalnum()) if DEFS else ""))
CSRC = os.path.join(ROOT, "self-supervised-normal-depth-estimation_b200", "csrc")


def build():
    deps = [SRC, os.path.join(ROOT, "include", "snde.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden",
                    "-I", os.path.join(ROOT, "include"), "-I", CSRC] + DEFS + ["-o", OUT, SRC], check=True)
    return OUT


def _img(a):   # numpy planar (n,3,h,w) any strides
    from snde_b200 import _abi
    return _abi.image_view(a.ctypes.data, [s // 4 for s in a.strides])


def _map(a):   # numpy (n,h,w)
    from snde_b200 import _abi
    return _abi.map_view(a.ctypes.data, [s // 4 for s in a.strides])


def run_loss(inp, alpha=0.85, w_rigid=1.0, w_smooth=0.5, flags=1, padding=0, masks=None, want_maps=False, seg_rows=0, fine_rows=0, fine_permille=0,
             want_image_grads=False):
    """inp: snde_b200.synthetic.LossInputs (CPU torch). Returns dict like oracle.compose.geonet_loss."""
    from snde_b200 import _abi
    lib = C.CDLL(build())
    lib.emul_sizeof_loss_args.restype = C.c_size_t
    assert lib.emul_sizeof_loss_args() == C.sizeof(_abi.SndeLossArgs)
    abi = _abi.lib()
    B, S, ns = inp.B, inp.S, inp.num_scales
    keep = []
    def tab(n):
        t = np.empty(n, np.float32)
        abi.snde_fill_grid_table(n, t.ctypes.data_as(_abi.c_float_p))
        keep.append(t)
        return t
    poses = np.ascontiguousarray(inp.poses.numpy(), np.float32)
    scales, outs = [], []
    for s in range(ns):
        tgt, src, depth = inp.tgt_pyr[s].numpy(), inp.src_pyr[s].numpy(), inp.depth_pyr[s].numpy()
        _, _, h, w = tgt.shape
        K = np.ascontiguousarray(inp.K_ms[:, s].numpy(), np.float32)
        gd = np.full(depth.shape, np.nan, np.float32)
        o = dict(grad_depth=gd)
        sc = dict(h=h, w=w, tgt=_img(tgt), src=_img(src), depth=_map(depth), grad_depth=_map(gd), K=K.ctypes.data,
                  xtab=tab(w).ctypes.data, ytab=tab(h).ctypes.data, photo_weight=w_rigid * S / 2, smooth_weight=w_smooth / 2 ** s)
        if masks is not None:
            mf, mb = masks[s]
            sc["mask_fwd"], sc["mask_bwd"] = _map(mf), _map(mb)
            keep += [mf, mb]
        if want_maps:
            for key in ("warped_fwd", "warped_bwd", "err_fwd", "err_bwd"):
                o[key] = np.full((B * S, 3, h, w), np.nan, np.float32)
                sc[key] = _img(o[key])
        if want_image_grads:
            o["grad_tgt_img"] = np.zeros(tgt.shape, np.float32); o["grad_src_img"] = np.zeros(src.shape, np.float32)
            sc["grad_tgt_img"], sc["grad_src_img"] = _img(o["grad_tgt_img"]), _img(o["grad_src_img"])
        keep += [tgt, src, depth, K]
        scales.append(sc); outs.append(o)
    loss_out = np.zeros(3 + 3 * ns, np.float32)
    gposes = np.full((B, S, 6), np.nan, np.float32)
    gK = np.full((B, ns, 3, 3), np.nan, np.float32)
    nbytes = abi.snde_loss_workspace_bytes(B, S, ns)
    ws = np.zeros(nbytes // 8 + 2, np.float64)
    wsp = (ws.ctypes.data + 15) & ~15
    args = _abi.make_loss_args(B, S, alpha, flags, padding, poses.ctypes.data, scales, loss_out.ctypes.data, gposes.ctypes.data,
                               gK.ctypes.data, wsp, nbytes)
    rc = lib.emul_loss_fwdbwd(C.byref(args), C.c_int(seg_rows), C.c_int(fine_rows), C.c_int(fine_permille))
    if rc < 0:
        raise RuntimeError(abi.snde_strerror(rc).decode())
    res = dict(variant=rc, loss_total=float(loss_out[0]), loss_rigid_warp=float(loss_out[1]), loss_disp_smooth=float(loss_out[2]),
               fwd_mean=loss_out[3::3].copy(), bwd_mean=loss_out[4::3].copy(), smooth=loss_out[5::3].copy(),
               grad_depth=[o["grad_depth"] for o in outs], grad_poses=gposes, grad_K_ms=gK)
    if want_maps:
        for key in ("warped_fwd", "warped_bwd", "err_fwd", "err_bwd"):
            res[key] = [o[key] for o in outs]
    if want_image_grads:
        for key in ("grad_tgt_img", "grad_src_img"):
            res[key] = [o[key] for o in outs]
    return res
