"""ctypes mirror of include/snde.h and the loader of libsnde.so.

There is NO fallback: if the CUDA library is missing or cannot be loaded, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNDE_LIB", os.path.join(HERE, "libsnde.so"))   # SNDE_LIB: tuning experiments only

SNDE_MAX_SCALES = 6
SNDE_PAD_ZEROS, SNDE_PAD_BORDER = 0, 1
SNDE_LOSS_GRAD, SNDE_LOSS_GRAD_K, SNDE_LOSS_NO_PHOTO, SNDE_LOSS_NO_SMOOTH, SNDE_LOSS_DISPARITY = 1, 2, 4, 8, 16

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class SndeImage(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sc", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


class SndeMap(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


SndeImageOut = SndeImage
SndeMapOut = SndeMap


class SndeLossScale(C.Structure):
    _fields_ = [
        ("h", C.c_int32), ("w", C.c_int32),
        ("tgt", SndeImage), ("src", SndeImage),
        ("depth", SndeMap), ("grad_depth", SndeMap),
        ("K", C.c_void_p), ("xtab", C.c_void_p), ("ytab", C.c_void_p),
        ("photo_weight", C.c_float), ("smooth_weight", C.c_float),
        ("mask_fwd", SndeMap), ("mask_bwd", SndeMap),
        ("warped_fwd", SndeImage), ("warped_bwd", SndeImage),
        ("err_fwd", SndeImage), ("err_bwd", SndeImage),
        ("grad_tgt_img", SndeImage), ("grad_src_img", SndeImage),
    ]


class SndeLossArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("S", C.c_int32), ("num_scales", C.c_int32),
        ("flags", C.c_int32), ("padding_mode", C.c_int32), ("alpha", C.c_float),
        ("poses", C.c_void_p),
        ("scale", SndeLossScale * SNDE_MAX_SCALES),
        ("loss_out", C.c_void_p), ("grad_poses", C.c_void_p), ("grad_K", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class SndeDnArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("rate", C.c_int32),
        ("thresh", C.c_float), ("clamp_max", C.c_float), ("w_depth", C.c_float), ("w_normal", C.c_float),
        ("depth", SndeMap), ("normal", SndeImage), ("xgrid", C.c_void_p), ("ygrid", C.c_void_p),
        ("normal_from_depth", SndeImage), ("depth_from_normal", SndeMap),
        ("loss_out", C.c_void_p), ("sq", C.c_void_p), ("sums", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/snde.h declares
SIGNATURES = {
    "snde_version": (C.c_int, []),
    "snde_strerror": (C.c_char_p, [C.c_int]),
    "snde_fill_grid_table": (C.c_int, [C.c_int, c_float_p]),
    "snde_sizeof_loss_args": (C.c_size_t, []),
    "snde_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "snde_loss_fwdbwd": (C.c_int, [C.POINTER(SndeLossArgs), C.c_void_p]),
    "snde_loss_num_tiles": (C.c_int64, [C.POINTER(SndeLossArgs)]),
    "snde_pose_matrices": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snde_pose_matrices_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "snde_rigid_flow_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int, SndeMap, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "snde_rigid_flow_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, SndeMap, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, SndeMap, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snde_bilinear_sampler_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, SndeImage, C.c_int, C.c_int, C.c_void_p,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snde_bilinear_sampler_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, SndeImage, C.c_int, C.c_int, C.c_void_p,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, SndeImage,
                                            C.c_void_p]),
    "snde_bilinear_sampler_bwd_deterministic": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, SndeImage, C.c_int, C.c_int, C.c_void_p,
                                                          C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, SndeImage,
                                                          C.c_void_p, C.c_void_p]),
    "snde_image_similarity_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, SndeImage, SndeImage, C.c_void_p, C.c_void_p]),
    "snde_image_similarity_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, SndeImage, SndeImage, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "snde_smooth_loss_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int, SndeMap, SndeImage, C.c_void_p, C.c_void_p]),
    "snde_smooth_loss_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, SndeMap, SndeImage, C.c_void_p, SndeMap, C.c_void_p]),
    "snde_frames_to_pyramid": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                         C.POINTER(C.c_void_p), C.c_void_p]),
    "snde_dn_consistency_fwd": (C.c_int, [C.POINTER(SndeDnArgs), C.c_void_p]),
    "snde_dn_consistency_bwd": (C.c_int, [C.POINTER(SndeDnArgs), C.c_void_p, SndeMap, C.c_void_p]),
    "snde_frames_to_pyramid_grouped": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                                 C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
}

_lib = None


class SndeLibraryError(RuntimeError):
    pass


def bind(lib: C.CDLL) -> C.CDLL:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def lib() -> C.CDLL:
    """The loaded CUDA library. Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SndeLibraryError(
                f"{LIB_PATH} not found: build it with `python -m snde_b200.build` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for these ops.")
        try:
            _lib = bind(C.CDLL(LIB_PATH))
            if _lib.snde_sizeof_loss_args() != C.sizeof(SndeLossArgs):
                raise SndeLibraryError("SndeLossArgs layout mismatch between _abi.py and libsnde.so")
        except OSError as e:  # e.g. libcudart missing
            raise SndeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(lib().snde_strerror(rc).decode())


# ----------------------------------------------------------------------------------------------
# argument marshalling shared by the torch wrapper and the (numpy, host-pointer) emulation tests
# ----------------------------------------------------------------------------------------------
def image_view(ptr, strides) -> SndeImage:
    """strides = element strides of the logical (n, c, y, x) axes"""
    return SndeImage(ptr, *[int(s) for s in strides])


def map_view(ptr, strides) -> SndeMap:
    """strides = element strides of the logical (n, y, x) axes"""
    return SndeMap(ptr, *[int(s) for s in strides])


def make_loss_args(B, S, alpha, flags, padding_mode, poses_ptr, scales, loss_out_ptr, grad_poses_ptr=None,
                   grad_K_ptr=None, workspace_ptr=None, workspace_bytes=0) -> SndeLossArgs:
    """scales: list of dicts with keys h, w, tgt, src (SndeImage), depth (SndeMap), K, xtab, ytab (pointers),
    photo_weight, smooth_weight and optional grad_depth, mask_fwd, mask_bwd (SndeMap), warped_fwd, warped_bwd,
    err_fwd, err_bwd (SndeImage)."""
    if len(scales) > SNDE_MAX_SCALES:
        raise ValueError(f"at most {SNDE_MAX_SCALES} scales")
    a = SndeLossArgs()
    a.B, a.S, a.num_scales, a.flags, a.padding_mode, a.alpha = B, S, len(scales), flags, padding_mode, alpha
    a.poses = poses_ptr
    for i, sc in enumerate(scales):
        d = a.scale[i]
        d.h, d.w = sc["h"], sc["w"]
        d.tgt, d.src, d.depth = sc["tgt"], sc["src"], sc["depth"]
        d.K, d.xtab, d.ytab = sc["K"], sc["xtab"], sc["ytab"]
        d.photo_weight, d.smooth_weight = sc["photo_weight"], sc["smooth_weight"]
        for key in ("grad_depth", "mask_fwd", "mask_bwd", "warped_fwd", "warped_bwd", "err_fwd", "err_bwd", "grad_tgt_img", "grad_src_img"):
            if sc.get(key) is not None:
                setattr(d, key, sc[key])
    a.loss_out, a.grad_poses, a.grad_K = loss_out_ptr, grad_poses_ptr, grad_K_ptr
    a.workspace, a.workspace_bytes = workspace_ptr, workspace_bytes
    return a
