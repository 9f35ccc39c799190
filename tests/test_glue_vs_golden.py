"""The caller-side glue the drop-in module restates in torch (scale_pyramid,
compute_multi_scale_intrinsics, euler2mat, pose_vec2mat, meshgrid, pixel2cam, cam2pixel) against known answers produced by
calling the reference's own functions (tests/golden/glue_kats.npz, make_golden.py::glue_kats).  CPU only: these are plain
torch ops (meshgrid's tables come from the C library's host helper snde_fill_grid_table)."""
import numpy as np
import pytest
import torch

from cases import load_golden


@pytest.fixture(scope="module")
def k():
    return load_golden("glue_kats")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_scale_pyramid_is_bit_identical_and_keeps_the_nhwc_view_layout(k):
    from snde_b200.utils_edited import scale_pyramid
    img = _t(k["sp_img"])
    levels = scale_pyramid(img, 4)
    assert len(levels) == 4
    for s, lvl in enumerate(levels):
        ref = k[f"sp_level_{s}"]
        assert tuple(lvl.shape) == ref.shape                         # (b, h_s, w_s, 3): logical NHWC
        assert lvl.permute(0, 3, 1, 2).is_contiguous()               # ... a VIEW of planar NCHW memory (reference :97, :106)
        assert np.array_equal(lvl.numpy(), ref)
    assert scale_pyramid(None, 4) is None
    # the synthetic generator's planar pyramid (the layout the fused kernel's fast path takes) holds the same numbers
    from snde_b200.synthetic import scale_pyramid as planar_pyramid
    for s, lvl in enumerate(planar_pyramid(img, 4)):
        assert np.array_equal(lvl.permute(0, 2, 3, 1).numpy(), k[f"sp_level_{s}"])


def test_compute_multi_scale_intrinsics_bit_identical(k):
    from snde_b200.synthetic import multi_scale_intrinsics
    from snde_b200.utils_edited import compute_multi_scale_intrinsics
    K = _t(k["msi_K"])
    assert np.array_equal(compute_multi_scale_intrinsics(K, 4).numpy(), k["msi_out"])
    assert np.array_equal(multi_scale_intrinsics(K, 4).numpy(), k["msi_out"])


def test_euler2mat_and_pose_vec2mat(k):
    from snde_b200.utils_edited import euler2mat, pose_vec2mat
    ang = _t(k["e2m_zyx"])
    got = euler2mat(ang[:, 0], ang[:, 1], ang[:, 2]).numpy()
    assert np.abs(got - k["e2m_out"]).max() <= 1.2e-7               # same products, the reference's bmm may contract differently
    kk = load_golden("kats")
    T = pose_vec2mat(_t(kk["pose_vec"])).numpy()
    assert np.abs(T - kk["pose_mat"]).max() <= 1.2e-7
    assert np.array_equal(T[:, 3], np.tile(np.array([0, 0, 0, 1], np.float32), (T.shape[0], 1)))


def test_meshgrid_bit_identical(k):
    from snde_b200.utils_edited import meshgrid
    for key in k.files:
        if not key.startswith("mg_"):
            continue
        hw, hom = key[3:].rsplit("_", 1)
        h, w = (int(v) for v in hw.split("x"))
        ref = k[key]
        got = meshgrid(ref.shape[0], h, w, bool(int(hom)), device=torch.device("cpu"))
        assert tuple(got.shape) == ref.shape
        assert np.array_equal(got.numpy(), ref)


def test_pixel2cam_and_cam2pixel(k):
    from snde_b200.utils_edited import cam2pixel, meshgrid, pixel2cam
    depth, K = _t(k["p2c_depth"]), _t(k["p2c_K"])
    b, h, w = depth.shape
    pix = meshgrid(b, h, w, True, device=torch.device("cpu"))
    cam = pixel2cam(depth, pix, K)
    assert tuple(cam.shape) == k["p2c_out"].shape
    assert np.array_equal(cam.numpy(), k["p2c_out"])
    assert np.array_equal(pixel2cam(depth, pix, K, False).numpy(), k["p2c_out_nohom"])
    out = cam2pixel(cam, _t(k["c2p_proj"]))
    assert tuple(out.shape) == k["c2p_out"].shape                     # (b, h, w, 2)
    assert np.array_equal(out.numpy(), k["c2p_out"])
