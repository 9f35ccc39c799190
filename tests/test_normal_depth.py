"""Depth <-> normal consistency (BASELINE config 3).  The reference's D2N / N2D (models/baseline.py:350-446,
495-547) are inference-only, with a scrambled patch extraction and no loss: PARITY IS UNPINNED BY THE REFERENCE.  The oracle
(oracle/normal_depth.py) restates the intended computation; it is pinned here by known answers (a plane is a fixed point of both
regressions), the CUDA kernels are then held to the oracle."""
import numpy as np
import pytest
import torch

from cases import rel_l2


def _plane(h, w, n, c, dtype=torch.float64):
    from oracle import normal_depth as ND
    g = ND._grid(h, w, 0.6, 0.4, dtype)
    n = torch.tensor(n, dtype=dtype)
    n = n / n.norm()
    depth = (c / (n.view(3, 1, 1) * g).sum(0)).unsqueeze(0)
    return depth, n.view(1, 3, 1, 1).expand(1, 3, h, w).contiguous()


def test_oracle_known_answers_plane_is_a_fixed_point():
    """points of a plane n . P = c: the least-squares normal of ANY window is n, and every neighbour's depth estimate of the centre
    is exact -- so D2N returns n, N2D returns the depth (up to the 1e-5 regulariser), the loss is ~0 and so is its gradient"""
    from oracle import normal_depth as ND
    depth, normal = _plane(48, 64, [0.2, -0.1, 0.97], 2.0)
    n2, d2 = ND.depth_to_normal(depth, normal), ND.normal_to_depth(depth, normal)
    assert (n2 - normal).abs().max().item() < 1e-10
    assert ((d2 - depth).abs() / depth).max().item() < 1e-5
    d = depth.clone().requires_grad_(True)
    loss, ld, ln = ND.consistency_loss(d, normal)
    assert loss.item() < 1e-5 and ln.item() < 1e-10
    # a bump on one pixel: N2D smooths it away, the loss gradient at that pixel pushes it back (positive: depth too large)
    dep = depth.clone(); dep[0, 24, 32] *= 1.05
    d = dep.clone().requires_grad_(True)
    loss, _, _ = ND.consistency_loss(d, normal)
    loss.backward()
    assert d.grad[0, 24, 32].item() > 0 and d.grad[0, 24, 32].item() == d.grad.max().item()


def test_oracle_window_is_spatial_and_zero_padded():
    """the window of pixel i is the dilated k x k spatial neighbourhood (the reference's F.unfold on an NHWC tensor is not);
    one neighbour switched off by the angle test leaves the plane's normal unchanged, a wrong depth there no longer matters"""
    from oracle import normal_depth as ND
    depth, normal = _plane(40, 56, [0.0, 0.0, 1.0], 1.5)
    normal = normal.clone(); depth = depth.clone()
    normal[0, :, 20, 28] = torch.tensor([1.0, 0.0, 0.0], dtype=normal.dtype)         # fails n_j . n_i > 0.95 for every neighbour
    depth[0, 20, 28] = 9.0
    n2 = ND.depth_to_normal(depth, normal)
    mask = torch.ones(40, 56, dtype=torch.bool); mask[20, 28] = False
    assert (n2[0][:, mask] - normal[0][:, mask]).abs().max().item() < 1e-9
    # ... and a pixel 4 (= rate) columns to the right sees it as the (0, -1) tap: switching the angle test off changes its normal
    n3 = ND.depth_to_normal(depth, normal, thresh=-2.0)
    assert (n3[0, :, 20, 32] - n2[0, :, 20, 32]).abs().max().item() > 1e-3
    assert (n3[0, :, 20, 33] - n2[0, :, 20, 33]).abs().max().item() < 1e-9           # not a tap of (20, 33)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 96, 128), (1, 61, 83), (3, 40, 200)])
def test_kernels_match_oracle(shape):
    from oracle import normal_depth as ND
    from snde_b200.normal_depth import depth_normal_consistency_loss, depth_normal_maps
    B, h, w = shape
    depth, normal = ND.synthetic_scene(B, h, w, seed=h)
    d, n = depth.cuda(), normal.cuda()
    n2, d2 = depth_normal_maps(d, n)
    on, od = ND.depth_to_normal(depth, normal), ND.normal_to_depth(depth, normal)
    on64 = ND.depth_to_normal(depth.double(), normal.double())
    # normals: the 3x3 normal equations are ill-conditioned where the window's valid points are nearly collinear; compare against
    # the fp64 solution and allow what the fp32 oracle itself loses there
    err = (n2.cpu().double() - on64).norm(dim=1)
    floor = (on.double() - on64).norm(dim=1)
    print(f"[d2n {shape}] max |n' - fp64| {err.max():.2e} (fp32 oracle: {floor.max():.2e}), median {err.median():.2e}")
    assert (err <= 5 * floor + 2e-4).all()
    assert rel_l2(d2.cpu().numpy(), od.numpy()) < 1e-5
    assert (d2.cpu() - od).abs().max().item() < 1e-4
    dd = d.clone().requires_grad_(True)
    loss, detail = depth_normal_consistency_loss(dd, n, w_depth=1.0, w_normal=0.5)
    (2.0 * loss).backward()
    od_ = depth.clone().requires_grad_(True)
    ol, old, oln = ND.consistency_loss(od_, normal, w_depth=1.0, w_normal=0.5)
    (2.0 * ol).backward()
    assert abs(detail[1].item() - old.item()) < 1e-5 * abs(old.item())
    assert abs(detail[2].item() - oln.item()) < 1e-4 * abs(oln.item()) + 1e-7
    assert abs(loss.item() - ol.item()) < 1e-5 * abs(ol.item())
    # a non-default window (k = 5, dilation 3) takes the generic kernel instance
    n5, d5 = depth_normal_maps(d, n, k=5, rate=3)
    assert rel_l2(d5.cpu().numpy(), ND.normal_to_depth(depth, normal, k=5, rate=3).numpy()) < 1e-5
    assert (n5.cpu().double() - ND.depth_to_normal(depth.double(), normal.double(), k=5, rate=3)).norm(dim=1).max().item() < 5e-3
    g, og = dd.grad.cpu().numpy(), od_.grad.numpy()
    # sign(d_i - d'_i) flips where the two agree to the last bits; ONE such pixel moves the gradient of the k*k pixels whose
    # window holds it, so up to three flipped signs are tolerated
    flips = np.abs(g - og) > 1e-3 * np.abs(og).max()
    print(f"[dn grad {shape}] rel {rel_l2(g[~flips], og[~flips]):.2e}, flipped {flips.mean():.2e}")
    assert rel_l2(g[~flips], og[~flips]) < 1e-4 and flips.sum() <= 3 * 81


@pytest.mark.gpu
def test_nyu_size_plane_fixed_point_and_strided_inputs():
    """BASELINE config 3 shape (480x640, B=8): a plane is a fixed point at full size; NHWC-strided normals give the same result"""
    from snde_b200.normal_depth import depth_normal_consistency_loss, depth_normal_maps
    depth, normal = _plane(480, 640, [0.15, 0.1, 0.98], 2.5, torch.float32)
    d, n = depth.expand(8, -1, -1).contiguous().cuda(), normal.expand(8, -1, -1, -1).contiguous().cuda()
    n2, d2 = depth_normal_maps(d, n)
    # at this resolution a window spans +-16 pixels = +-0.03 of the normalised grid: the points are nearly collinear seen from the
    # origin and the fp32 normal equations lose digits -- the fp32 oracle (LU solve) by as much as the kernel
    from oracle import normal_depth as ND
    floor = (ND.depth_to_normal(depth, normal) - normal).abs().max().item()
    err = (n2 - n).abs().max().item()
    print(f"[d2n 480x640 plane] max |n' - n|: kernel {err:.2e}, fp32 oracle {floor:.2e}")
    assert err < 3 * floor + 1e-3
    assert ((d2 - d).abs() / d).max().item() < 1e-4
    loss, detail = depth_normal_consistency_loss(d, n)
    assert detail[1].item() < 1e-4 and detail[2].item() < 1e-5
    n_nhwc = n.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    n3, d3 = depth_normal_maps(d, n_nhwc)
    assert torch.equal(n3, n2) and torch.equal(d3, d2)
