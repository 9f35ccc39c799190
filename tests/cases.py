"""Generator arguments of the committed golden cases (must match tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "small_smooth": dict(B=2, S=2, H=32, W=64, num_scales=3, seed=1, smooth=True, motion="small"),
    "small_noise_large_motion": dict(B=2, S=2, H=24, W=40, num_scales=2, seed=2, smooth=False, motion="large"),
    "small_s1_nyu": dict(B=1, S=1, H=30, W=44, num_scales=2, seed=3, smooth=True, motion="large", intrinsics="nyu"),
    "config1_noise": dict(B=4, S=2, H=128, W=416, num_scales=4, seed=0, smooth=False, motion="small"),
    "config1_smooth": dict(B=4, S=2, H=128, W=416, num_scales=4, seed=0, smooth=True, motion="small"),
}
SMALL = ["small_smooth", "small_noise_large_motion", "small_s1_nyu"]


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


def nhwc(t):
    """planar torch (n,3,h,w) -> contiguous numpy NHWC"""
    return np.ascontiguousarray(t.permute(0, 2, 3, 1).numpy())


def oracle_inputs(inp):
    """LossInputs (torch, planar) -> numpy arguments of oracle.compose.geonet_loss"""
    return dict(tgt_pyr=[nhwc(t) for t in inp.tgt_pyr], src_pyr=[nhwc(t) for t in inp.src_pyr],
                depth_pyr=[d.numpy() for d in inp.depth_pyr], poses=inp.poses.numpy(), K_ms=inp.K_ms.numpy())


def close_except_flips(a, b, rtol=1e-4, max_flip_frac=1e-4, flip_level=1e-3):
    """Gradient maps agree to `rtol` (relative L2) once isolated DISCONTINUITY pixels are set aside.

    d(warped)/d(coords) of a bilinear sampler jumps when a coordinate crosses an integer, sign(x - y) and the
    clamp are step functions: ~1e-5 px of legitimate fp32 difference in the coordinates (closed-form vs LU
    matrix inverse, summation order) flips a handful of pixels per million (hard part 2).
    Returns (ok, rel_l2_without_flips, flip_fraction)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    flips = d > flip_level * np.abs(b).max()
    frac = flips.mean()
    keep = ~flips
    rel = float(np.sqrt((d[keep] ** 2).sum()) / max(np.sqrt((b[keep] ** 2).sum()), 1e-30))
    return (rel < rtol and frac <= max_flip_frac), rel, float(frac)
