"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

ctypes/numpy binding of ``oracle/snde_oracle.c`` (a plain-C restatement of the
reference's rigid-warp + photometric + smoothness loss path) plus the loss
composition of ``GeoNetModel.build_rigid_warp_flow`` / ``build_losses``
(reference ``models/baseline.py:867-986, 1115-1161``) in :mod:`oracle.compose`.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``snde_b200``) never does: it fails loudly when its CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors of its own; this
oracle is pinned against the reference executed in the authoring container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force: bool = False) -> None:
    """Compile the C oracle (gcc; both precisions) into oracle/_build/."""
    libs = [os.path.join(_BUILD, f"libsnde_oracle_{p}.so") for p in ("f32", "f64")]
    src = os.path.join(_HERE, "snde_oracle.c")
    fresh = all(os.path.exists(l) and os.path.getmtime(l) >= os.path.getmtime(src) for l in libs)
    if fresh and not force:
        return
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Oracle:
    """One precision of the C oracle. ``Oracle('f32')`` follows the reference's fp32
    arithmetic; ``Oracle('f64')`` is the same algorithm in double (noise-floor probe)."""

    def __init__(self, precision: str = "f32"):
        assert precision in ("f32", "f64")
        build()
        self.lib = ctypes.CDLL(os.path.join(_BUILD, f"libsnde_oracle_{precision}.so"))
        self.dt = np.float32 if precision == "f32" else np.float64
        self.creal = ctypes.c_float if precision == "f32" else ctypes.c_double
        assert self.lib.snde_oracle_real_bytes() == np.dtype(self.dt).itemsize
        self.lib.snde_oracle_smooth_loss.restype = ctypes.c_double
        self.lib.snde_oracle_mean.restype = ctypes.c_double

    def _a(self, x):
        return np.ascontiguousarray(np.asarray(x), dtype=self.dt)

    # -- geometry -------------------------------------------------------------------
    def grid_table(self, n: int) -> np.ndarray:
        out = np.empty(n, self.dt)
        self.lib.snde_oracle_grid_table(ctypes.c_int(n), _p(out))
        return out

    def pose_vec2mat(self, vec) -> np.ndarray:
        vec = self._a(vec)
        out = np.empty((vec.shape[0], 4, 4), self.dt)
        for i in range(vec.shape[0]):
            self.lib.snde_oracle_pose_vec2mat(_p(vec[i]), _p(out[i]))
        return out

    def inverse(self, A) -> np.ndarray:
        A = self._a(A)
        n = A.shape[-1]
        out = np.empty_like(A)
        flatA, flatO = A.reshape(-1, n, n), out.reshape(-1, n, n)
        for i in range(flatA.shape[0]):
            rc = self.lib.snde_oracle_inverse(ctypes.c_int(n), _p(flatA[i]), _p(flatO[i]))
            if rc:
                raise np.linalg.LinAlgError("singular")
        return out

    def pose_prep(self, poses, K):
        poses, K = self._a(poses), self._a(K)
        B, S = poses.shape[:2]
        Pf = np.empty((S * B, 3, 4), self.dt)
        Pb = np.empty((S * B, 3, 4), self.dt)
        Kinv = np.empty((B, 3, 3), self.dt)
        rc = self.lib.snde_oracle_pose_prep(B, S, _p(poses), _p(K), _p(Pf), _p(Pb), _p(Kinv))
        if rc:
            raise np.linalg.LinAlgError("singular")
        return Pf, Pb, Kinv

    def pose_prep_bwd(self, poses, K, gPf, gPb, gKinv=None):
        poses, K = self._a(poses), self._a(K)
        B, S = poses.shape[:2]
        gPf = np.ascontiguousarray(gPf, np.float64)
        gPb = np.ascontiguousarray(gPb, np.float64)
        gKinv = None if gKinv is None else np.ascontiguousarray(gKinv, np.float64)
        gposes = np.zeros((B, S, 6), np.float64)
        gK = np.zeros((B, 3, 3), np.float64)
        rc = self.lib.snde_oracle_pose_prep_bwd(B, S, _p(poses), _p(K), _p(gPf), _p(gPb), _p(gKinv), _p(gposes), _p(gK))
        if rc:
            raise np.linalg.LinAlgError("singular")
        return gposes, gK

    def rigid_flow(self, depth, Kinv, P):
        depth, Kinv, P = self._a(depth), self._a(Kinv), self._a(P)
        B, h, w = depth.shape
        flow = np.empty((B, h, w, 2), self.dt)
        self.lib.snde_oracle_rigid_flow(B, h, w, _p(depth), _p(Kinv), _p(P), _p(flow))
        return flow

    def rigid_flow_bwd(self, depth, Kinv, P, gflow, want_gKinv=True):
        depth, Kinv, P, gflow = self._a(depth), self._a(Kinv), self._a(P), self._a(gflow)
        B, h, w = depth.shape
        gdepth = np.empty((B, h, w), self.dt)
        gP = np.zeros((B, 3, 4), np.float64)
        gKinv = np.zeros((B, 3, 3), np.float64) if want_gKinv else None
        self.lib.snde_oracle_rigid_flow_bwd(B, h, w, _p(depth), _p(Kinv), _p(P), _p(gflow), _p(gdepth), _p(gP), _p(gKinv))
        return gdepth, gP, gKinv

    def flow_to_coords(self, flow):
        flow = self._a(flow)
        N, h, w, _ = flow.shape
        out = np.empty_like(flow)
        self.lib.snde_oracle_flow_to_coords(N, h, w, _p(flow), _p(out))
        return out

    # -- sampling -------------------------------------------------------------------
    def bilinear_sampler(self, img, coords):
        img, coords = self._a(img), self._a(coords)
        N, hs, ws, C = img.shape
        _, ht, wt, _ = coords.shape
        out = np.empty((N, ht, wt, C), self.dt)
        self.lib.snde_oracle_bilinear_sampler(N, hs, ws, C, _p(img), ht, wt, _p(coords), _p(out))
        return out

    def bilinear_sampler_bwd(self, img, coords, gout, want_gimg=False):
        img, coords, gout = self._a(img), self._a(coords), self._a(gout)
        N, hs, ws, C = img.shape
        _, ht, wt, _ = coords.shape
        gcoords = np.empty((N, ht, wt, 2), self.dt)
        gimg = np.zeros_like(img) if want_gimg else None
        self.lib.snde_oracle_bilinear_sampler_bwd(N, hs, ws, C, _p(img), ht, wt, _p(coords), _p(gout), _p(gcoords), _p(gimg))
        return gcoords, gimg

    # -- photometric ----------------------------------------------------------------
    def image_similarity(self, alpha, x, y):
        x, y = self._a(x), self._a(y)
        N, h, w, C = x.shape
        err = np.empty_like(x)
        self.lib.snde_oracle_image_similarity(N, h, w, C, self.creal(alpha), _p(x), _p(y), _p(err))
        return err

    def image_similarity_bwd(self, alpha, x, y, gerr):
        x, y, gerr = self._a(x), self._a(y), self._a(gerr)
        N, h, w, C = x.shape
        gx, gy = np.empty_like(x), np.empty_like(x)
        self.lib.snde_oracle_image_similarity_bwd(N, h, w, C, self.creal(alpha), _p(x), _p(y), _p(gerr), _p(gx), _p(gy))
        return gx, gy

    # -- smoothness -----------------------------------------------------------------
    def smooth_loss(self, depth, img) -> float:
        depth, img = self._a(depth), self._a(img)
        M, h, w, C = img.shape
        depth = depth.reshape(M, h, w)
        return float(self.lib.snde_oracle_smooth_loss(M, h, w, C, _p(depth), _p(img)))

    def smooth_loss_bwd(self, depth, img, gscalar=1.0):
        depth, img = self._a(depth), self._a(img)
        M, h, w, C = img.shape
        depth = depth.reshape(M, h, w)
        g = np.zeros((M, h, w), self.dt)
        self.lib.snde_oracle_smooth_loss_bwd(M, h, w, C, _p(depth), _p(img), self.creal(gscalar), _p(g))
        return g

    def mean(self, v) -> float:
        v = self._a(v)
        return float(self.lib.snde_oracle_mean(ctypes.c_size_t(v.size), _p(v)))
