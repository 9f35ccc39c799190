"""torch <-> C-ABI glue: every function here marshals CUDA tensors into include/snde.h calls on the
current stream.  PyTorch is plumbing (device memory, streams, autograd graph); the arithmetic is in
libsnde.so.  There is no CPU path: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi

_tables: Dict[Tuple[int, str], torch.Tensor] = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("snde_b200 ops need CUDA tensors (there is no CPU fallback)")
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("snde_b200 ops are fp32")


def grid_table(n: int, device) -> torch.Tensor:
    """Device copy of the reference's meshgrid table (utils/utils_edited.py:307-317), cached."""
    key = (n, str(device))
    t = _tables.get(key)
    if t is None:
        host = np.empty(n, np.float32)
        _abi.check(_abi.lib().snde_fill_grid_table(n, host.ctypes.data_as(_abi.c_float_p)))
        t = torch.from_numpy(host).to(device)
        _tables[key] = t
    return t


def img_nchw(t: torch.Tensor) -> _abi.SndeImage:
    """tensor shaped (n,3,h,w), any strides"""
    return _abi.image_view(t.data_ptr(), t.stride())


def img_nhwc(t: torch.Tensor) -> _abi.SndeImage:
    """tensor shaped (n,h,w,c), any strides (the reference's logical layout)"""
    s = t.stride()
    return _abi.image_view(t.data_ptr(), (s[0], s[3], s[1], s[2]))


def map_nhw(t: torch.Tensor) -> _abi.SndeMap:
    return _abi.map_view(t.data_ptr(), t.stride())


# ------------------------------------------------------------------------------------------------
# fused loss
# ------------------------------------------------------------------------------------------------
class FusedLossPlan:
    """Pre-marshalled arguments + output/workspace buffers of one fused-loss call SHAPE.

    Two ways to use it:
      * static (CUDA-graph style, bench.py, tests): construct once on fixed input tensors, ``run()`` any number of times;
        outputs live in ``loss_out``, ``grad_depth``, ``grad_poses`` (rewritten by every run).  The plan keeps the input
        tensors alive in this mode only.
      * rebinding (the autograd path, a trainer whose depth / poses are fresh network outputs every step): ``rebind(...)``
        patches the pointer fields of the cached ``SndeLossArgs`` for tensors of the SAME shapes and strides -- a few ctypes
        stores, no allocation, no marshalling -- and ``fresh_outputs()`` points the call at newly allocated gradient /
        loss tensors that the caller then owns.  Nothing of the caller's is retained.
    """

    def __init__(self, tgt_pyr: Sequence[torch.Tensor], src_pyr: Sequence[torch.Tensor], depth_pyr: Sequence[torch.Tensor],
                 poses: torch.Tensor, K_ms: torch.Tensor, alpha: float = 0.85, photo_weights: Optional[Sequence[float]] = None,
                 smooth_weights: Optional[Sequence[float]] = None, want_grad: bool = True, want_grad_K: bool = False,
                 padding_mode: str = "zeros", masks: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]] = None,
                 want_maps: bool = False, no_photo: bool = False, no_smooth: bool = False, disparity: bool = False,
                 keep_inputs: bool = True, want_image_grads: bool = False):
        ns = len(tgt_pyr)
        _require_cuda(poses, K_ms, *tgt_pyr, *src_pyr, *depth_pyr)
        B, S = poses.shape[0], poses.shape[1]
        dev = poses.device
        if poses.shape != (B, S, 6) or K_ms.shape != (B, ns, 3, 3):
            raise ValueError("poses must be (B,S,6) and K_ms (B,scales,3,3)")
        self.B, self.S, self.ns, self.device = B, S, ns, dev
        self.want_grad, self.want_grad_K = want_grad, want_grad_K
        self.want_image_grads = bool(want_image_grads)
        if want_image_grads and not want_grad:
            raise ValueError("want_image_grads needs want_grad")
        # d loss_rigid_warp / d image (section 8(f3)): ACCUMULATED by the kernel (atomics), so they are zero-filled before every run
        self.grad_tgt_img = [torch.zeros(t.shape, device=dev) for t in tgt_pyr] if want_image_grads else []
        self.grad_src_img = [torch.zeros(t.shape, device=dev) for t in src_pyr] if want_image_grads else []
        self._img_shapes = [(tuple(t.shape), tuple(u.shape)) for t, u in zip(tgt_pyr, src_pyr)]
        # per-scale intrinsics as contiguous (B,3,3) blocks (what the C call points at): one scale-major buffer
        self.K_buf = torch.empty(ns, B, 3, 3, device=dev)
        self.K = [self.K_buf[s] for s in range(ns)]
        if photo_weights is None:
            photo_weights = [1.0 * S / 2] * ns            # models/baseline.py:1126-1129 with loss_weight_rigid_warp = 1
        if smooth_weights is None:
            smooth_weights = [0.5 / 2 ** s for s in range(ns)]  # :1133 with loss_weight_disparity_smooth = 0.5
        self.loss_out = torch.zeros(3 + 3 * ns, device=dev)
        self.grad_depth = [torch.empty_like(d, memory_format=torch.contiguous_format) for d in depth_pyr] if want_grad else []
        self.grad_poses = torch.zeros(B, S, 6, device=dev) if want_grad else None
        self.grad_K = torch.zeros(B, ns, 3, 3, device=dev) if want_grad_K else None
        nbytes = _abi.lib().snde_loss_workspace_bytes(B, S, ns)
        self.workspace = torch.empty(nbytes // 8 + 2, dtype=torch.float64, device=dev)
        wsp = (self.workspace.data_ptr() + 15) & ~15
        self.maps = []
        self._depth_shapes = [tuple(d.shape) for d in depth_pyr]
        scales = []
        for s in range(ns):
            tgt, src, depth = tgt_pyr[s], src_pyr[s], depth_pyr[s]
            _, c, h, w = tgt.shape
            if c != 3 or src.shape != (B * S, 3, h, w) or depth.shape != (B * (1 + S), h, w):
                raise ValueError(f"scale {s}: expected tgt (B,3,h,w), src (S*B,3,h,w), depth (B*(1+S),h,w)")
            sc = dict(h=h, w=w, tgt=img_nchw(tgt), src=img_nchw(src), depth=map_nhw(depth), K=self.K[s].data_ptr(),
                      xtab=grid_table(w, dev).data_ptr(), ytab=grid_table(h, dev).data_ptr(),
                      photo_weight=float(photo_weights[s]), smooth_weight=float(smooth_weights[s]))
            if want_grad:
                sc["grad_depth"] = map_nhw(self.grad_depth[s])
            if masks is not None:
                mf, mb = masks[s]
                _require_cuda(mf, mb)
                sc["mask_fwd"], sc["mask_bwd"] = map_nhw(mf), map_nhw(mb)
            if want_image_grads:
                sc["grad_tgt_img"], sc["grad_src_img"] = img_nchw(self.grad_tgt_img[s]), img_nchw(self.grad_src_img[s])
            if want_maps:
                m = {k: torch.empty(B * S, 3, h, w, device=dev) for k in ("warped_fwd", "warped_bwd", "err_fwd", "err_bwd")}
                for k, v in m.items():
                    sc[k] = img_nchw(v)
                self.maps.append(m)
            scales.append(sc)
        flags = (_abi.SNDE_LOSS_GRAD if want_grad else 0) | (_abi.SNDE_LOSS_GRAD_K if want_grad_K else 0) \
            | (_abi.SNDE_LOSS_NO_PHOTO if no_photo else 0) | (_abi.SNDE_LOSS_NO_SMOOTH if no_smooth else 0) \
            | (_abi.SNDE_LOSS_DISPARITY if disparity else 0)
        pad = {"zeros": _abi.SNDE_PAD_ZEROS, "border": _abi.SNDE_PAD_BORDER}[padding_mode]
        self.poses = poses.contiguous()
        self.args = _abi.make_loss_args(B, S, float(alpha), flags, pad, self.poses.data_ptr(), scales, self.loss_out.data_ptr(),
                                        self.grad_poses.data_ptr() if want_grad else None,
                                        self.grad_K.data_ptr() if want_grad_K else None, wsp, nbytes)
        self._fn = _abi.lib().snde_loss_fwdbwd
        self._ref = C.byref(self.args)
        self._num_tiles = None
        self.kernel_launches = 3   # prep (also zeroes the accumulators), strip, finalize
        self._set_intrinsics(K_ms)
        self._keep = [tgt_pyr, src_pyr, depth_pyr, masks] if keep_inputs else None
        if not keep_inputs:
            self.poses = None

    @property
    def num_tiles(self) -> int:
        """strip tasks of one call (host-side preview of the library's segment-height rule)"""
        if self._num_tiles is None:
            self._num_tiles = int(_abi.lib().snde_loss_num_tiles(self._ref))
        return self._num_tiles

    def _set_intrinsics(self, K_ms: torch.Tensor) -> None:
        if K_ms.data_ptr() == self.K_buf.data_ptr() and K_ms.stride() == (9, self.B * 9, 3, 1):
            return                                        # the caller already filled this very buffer (a permuted view of it)
        self.K_buf.copy_(K_ms.permute(1, 0, 2, 3))        # one small strided copy

    def refresh_intrinsics(self, K_ms: torch.Tensor) -> None:
        """Re-copy the per-scale intrinsics (the plan keeps a contiguous scale-major copy of the strided K_ms[:, s])."""
        self._set_intrinsics(K_ms)

    def rebind(self, tgt_pyr, src_pyr, depth_pyr, poses: torch.Tensor, K_ms: torch.Tensor, masks=None) -> None:
        """Point the cached arguments at other tensors of the same shapes AND strides (the caller checks that: the plan is
        looked up by them).  poses must be contiguous."""
        a = self.args
        a.poses = poses.data_ptr()
        for s in range(self.ns):
            sc = a.scale[s]
            sc.tgt.ptr, sc.src.ptr, sc.depth.ptr = tgt_pyr[s].data_ptr(), src_pyr[s].data_ptr(), depth_pyr[s].data_ptr()
            if masks is not None:
                sc.mask_fwd.ptr, sc.mask_bwd.ptr = masks[s][0].data_ptr(), masks[s][1].data_ptr()
        self._set_intrinsics(K_ms)

    def fresh_outputs(self):
        """Newly allocated loss / gradient tensors for the next run (owned by the caller afterwards).
        Returns (loss_out, grad_poses | None, grad_K | None, [grad_depth])."""
        a, dev = self.args, self.device
        loss_out = torch.empty(3 + 3 * self.ns, device=dev)
        a.loss_out = loss_out.data_ptr()
        gposes = gK = None
        gdepth = []
        if self.want_grad:
            gposes = torch.empty(self.B, self.S, 6, device=dev)
            a.grad_poses = gposes.data_ptr()
            for s, shp in enumerate(self._depth_shapes):
                g = torch.empty(shp, device=dev)
                a.scale[s].grad_depth.ptr = g.data_ptr()
                gdepth.append(g)
        if self.want_grad_K:
            gK = torch.empty(self.B, self.ns, 3, 3, device=dev)
            a.grad_K = gK.data_ptr()
        if self.want_image_grads:
            self.grad_tgt_img = [torch.zeros(t, device=dev) for t, _ in self._img_shapes]
            self.grad_src_img = [torch.zeros(u, device=dev) for _, u in self._img_shapes]
            for s in range(self.ns):
                a.scale[s].grad_tgt_img.ptr, a.scale[s].grad_src_img.ptr = self.grad_tgt_img[s].data_ptr(), self.grad_src_img[s].data_ptr()
            self._fresh_img = True
        return loss_out, gposes, gK, gdepth

    def run(self) -> None:
        """Launch on the current stream (asynchronous)."""
        if self.want_image_grads:
            if getattr(self, "_fresh_img", False):
                self._fresh_img = False                   # fresh_outputs() just handed out zero-filled buffers
            else:
                torch._foreach_zero_(self.grad_tgt_img + self.grad_src_img)
        rc = self._fn(self._ref, _stream())
        if rc:
            _abi.check(rc)


# ------------------------------------------------------------------------------------------------
# reference-signature pieces
# ------------------------------------------------------------------------------------------------
def pose_matrices(pose: torch.Tensor, K: torch.Tensor, reverse: bool):
    _require_cuda(pose, K)
    B = pose.shape[0]
    pose = pose if pose.stride(1) == 1 else pose.contiguous()
    K = K.contiguous()
    P = torch.empty(B, 3, 4, device=pose.device)
    Kinv = torch.empty(B, 3, 3, device=pose.device)
    _abi.check(_abi.lib().snde_pose_matrices(B, pose.data_ptr(), pose.stride(0), K.data_ptr(), int(reverse), P.data_ptr(),
                                              Kinv.data_ptr(), _stream()))
    return P, Kinv


def pose_matrices_bwd(pose, K, reverse, gP, gKinv, want_gK):
    B = pose.shape[0]
    pose = pose if pose.stride(1) == 1 else pose.contiguous()
    K = K.contiguous()
    gpose = torch.empty(B, 6, device=pose.device)
    gK = torch.empty(B, 3, 3, device=pose.device) if want_gK else None
    _abi.check(_abi.lib().snde_pose_matrices_bwd(B, pose.data_ptr(), pose.stride(0), K.data_ptr(), int(reverse), gP.data_ptr(),
                                                  gKinv.data_ptr() if gKinv is not None else None, gpose.data_ptr(),
                                                  gK.data_ptr() if want_gK else None, _stream()))
    return gpose, gK


def rigid_flow_fwd(depth, Kinv, P):
    _require_cuda(depth, Kinv, P)
    B, h, w = depth.shape
    flow = torch.empty(B, h, w, 2, device=depth.device)
    dev = depth.device
    _abi.check(_abi.lib().snde_rigid_flow_fwd(B, h, w, map_nhw(depth), Kinv.data_ptr(), P.data_ptr(), grid_table(w, dev).data_ptr(),
                                               grid_table(h, dev).data_ptr(), flow.data_ptr(), _stream()))
    return flow


def rigid_flow_bwd(depth, Kinv, P, gflow, want_gKinv):
    B, h, w = depth.shape
    dev = depth.device
    gflow = gflow.contiguous()
    gdepth = torch.empty(B, h, w, device=dev)
    gP = torch.zeros(B, 3, 4, dtype=torch.float64, device=dev)
    gKinv = torch.zeros(B, 3, 3, dtype=torch.float64, device=dev) if want_gKinv else None
    _abi.check(_abi.lib().snde_rigid_flow_bwd(B, h, w, map_nhw(depth), Kinv.data_ptr(), P.data_ptr(), grid_table(w, dev).data_ptr(),
                                               grid_table(h, dev).data_ptr(), gflow.data_ptr(), map_nhw(gdepth), gP.data_ptr(),
                                               gKinv.data_ptr() if want_gKinv else None, _stream()))
    return gdepth, gP, gKinv


def _pad(mode: str) -> int:
    return {"zeros": _abi.SNDE_PAD_ZEROS, "border": _abi.SNDE_PAD_BORDER}[mode]


def sampler_fwd(imgs_nhwc, coords, add_grid: bool, padding_mode: str, want_valid: bool):
    _require_cuda(imgs_nhwc, coords)
    N, hs, ws, Cc = imgs_nhwc.shape
    _, ht, wt, _ = coords.shape
    dev = imgs_nhwc.device
    coords = coords.contiguous()
    out = torch.empty(N, ht, wt, Cc, device=dev)
    valid = torch.empty(N, ht, wt, device=dev) if want_valid else None
    xt = grid_table(wt, dev).data_ptr() if add_grid else None
    yt = grid_table(ht, dev).data_ptr() if add_grid else None
    _abi.check(_abi.lib().snde_bilinear_sampler_fwd(N, Cc, hs, ws, img_nhwc(imgs_nhwc), ht, wt, coords.data_ptr(), int(add_grid),
                                                     xt, yt, _pad(padding_mode), out.data_ptr(),
                                                     valid.data_ptr() if want_valid else None, _stream()))
    return out, valid


# image-gradient scatter mode of bilinear_sampler / flow_warp / inverse_warp backward: False = float atomics (fast, last bits vary
# from run to run), True = the deterministic two-pass mode (64-bit fixed-point accumulation, bit-identical every run)
DETERMINISTIC_IMAGE_GRADIENT = False


def sampler_bwd(imgs_nhwc, coords, add_grid: bool, padding_mode: str, gout, want_gimg: bool):
    N, hs, ws, Cc = imgs_nhwc.shape
    _, ht, wt, _ = coords.shape
    dev = imgs_nhwc.device
    coords, gout = coords.contiguous(), gout.contiguous()
    gcoords = torch.empty(N, ht, wt, 2, device=dev)
    gimg = torch.zeros(N, hs, ws, Cc, device=dev) if want_gimg else None
    xt = grid_table(wt, dev).data_ptr() if add_grid else None
    yt = grid_table(ht, dev).data_ptr() if add_grid else None
    gview = img_nhwc(gimg) if want_gimg else _abi.SndeImage()
    if want_gimg and DETERMINISTIC_IMAGE_GRADIENT:
        scratch = torch.empty(N * Cc * hs * ws + 1, dtype=torch.int64, device=dev)
        _abi.check(_abi.lib().snde_bilinear_sampler_bwd_deterministic(N, Cc, hs, ws, img_nhwc(imgs_nhwc), ht, wt, coords.data_ptr(),
                                                                       int(add_grid), xt, yt, _pad(padding_mode), gout.data_ptr(),
                                                                       gcoords.data_ptr(), gview, scratch.data_ptr(), _stream()))
        return gcoords, gimg
    _abi.check(_abi.lib().snde_bilinear_sampler_bwd(N, Cc, hs, ws, img_nhwc(imgs_nhwc), ht, wt, coords.data_ptr(), int(add_grid),
                                                     xt, yt, _pad(padding_mode), gout.data_ptr(), gcoords.data_ptr(), gview,
                                                     _stream()))
    return gcoords, gimg


def similarity_fwd(alpha: float, x_nhwc, y_nhwc):
    _require_cuda(x_nhwc, y_nhwc)
    N, h, w, Cc = x_nhwc.shape
    if Cc != 3 or y_nhwc.shape != x_nhwc.shape:
        raise ValueError("image_similarity expects two (N,h,w,3) tensors")
    err = torch.empty(N, h, w, 3, device=x_nhwc.device)
    _abi.check(_abi.lib().snde_image_similarity_fwd(N, h, w, float(alpha), img_nhwc(x_nhwc), img_nhwc(y_nhwc), err.data_ptr(), _stream()))
    return err


def similarity_bwd(alpha: float, x_nhwc, y_nhwc, gerr, want_gx: bool, want_gy: bool):
    N, h, w, _ = x_nhwc.shape
    gerr = gerr.contiguous()
    gx = torch.empty(N, h, w, 3, device=gerr.device) if want_gx else None
    gy = torch.empty(N, h, w, 3, device=gerr.device) if want_gy else None
    _abi.check(_abi.lib().snde_image_similarity_bwd(N, h, w, float(alpha), img_nhwc(x_nhwc), img_nhwc(y_nhwc), gerr.data_ptr(),
                                                     gx.data_ptr() if want_gx else None, gy.data_ptr() if want_gy else None, _stream()))
    return gx, gy


def smooth_fwd(depth_mhw, img_nhwc_):
    _require_cuda(depth_mhw, img_nhwc_)
    M, h, w = depth_mhw.shape
    sums = torch.zeros(2, dtype=torch.float64, device=depth_mhw.device)
    _abi.check(_abi.lib().snde_smooth_loss_fwd(M, h, w, map_nhw(depth_mhw), img_nhwc(img_nhwc_), sums.data_ptr(), _stream()))
    return (sums[0] / float(M * h * (w - 1)) + sums[1] / float(M * (h - 1) * w)).float()


def smooth_bwd(depth_mhw, img_nhwc_, gscale):
    M, h, w = depth_mhw.shape
    g = torch.empty(M, h, w, device=depth_mhw.device)
    gscale = gscale.reshape(1).contiguous().float()
    _abi.check(_abi.lib().snde_smooth_loss_bwd(M, h, w, map_nhw(depth_mhw), img_nhwc(img_nhwc_), gscale.data_ptr(), map_nhw(g), _stream()))
    return g


def frames_to_pyramid(frames_u8: torch.Tensor, num_scales: int, out: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
    """uint8 frames (n,3*G,H,W), any strides -> list of planar fp32 levels (G*n,3,H>>s,W>>s) with values (u8/255)*2-1 averaged
    over 2^s x 2^s blocks: the image part of iter_data_preparation fused with scale_pyramid (ONE kernel launch).
    G > 1 is the DataLoader's source block (B,3*S,H,W): group g of frame b lands in image g*n + b (source-major,
    models/baseline.py:772-775).  out: optional pre-allocated contiguous levels."""
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8:
        raise RuntimeError("frames_to_pyramid needs a CUDA uint8 tensor (there is no CPU fallback)")
    n, c, H, W = frames_u8.shape
    if c % 3 or c == 0:
        raise ValueError("frames must be (n,3*G,H,W)")
    if not 1 <= num_scales <= _abi.SNDE_MAX_SCALES:
        raise ValueError(f"num_scales must be 1..{_abi.SNDE_MAX_SCALES}")
    R = 1 << (num_scales - 1)
    if H % R or W % R:
        raise RuntimeError(f"frames_to_pyramid: H and W must be multiples of 2^(num_scales-1) = {R} (got {H}x{W})")
    G = c // 3
    levels = list(out) if out is not None else [torch.empty(G * n, 3, H >> s, W >> s, device=frames_u8.device) for s in range(num_scales)]
    for s, t in enumerate(levels):
        if t.shape != (G * n, 3, H >> s, W >> s) or not t.is_contiguous() or t.dtype != torch.float32:
            raise ValueError("pyramid level buffers must be contiguous fp32 (G*n,3,H>>s,W>>s)")
    ptrs = (C.c_void_p * num_scales)(*[t.data_ptr() for t in levels])
    st = frames_u8.stride()
    _abi.check(_abi.lib().snde_frames_to_pyramid_grouped(n, G, H, W, num_scales, frames_u8.data_ptr(), st[0], st[1], st[2], st[3], ptrs, _stream()))
    return levels
