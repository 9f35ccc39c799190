"""The fused loss STRIP PROGRAM (csrc/snde_strip.cuh), compiled for the host by tests/host_emul (32 fibers per
warp) and run on numpy buffers, against the oracle.  This checks the device code's logic (task table, halo columns,
rolling row sums, segment seams, zero padding, gradient ownership, finalize) in the GPU-less container; the GPU
tests then check the real kernels."""
import numpy as np
import pytest

import emul
from cases import CASES, EDGE, SMALL, close_except_flips, load_golden, oracle_inputs, rel_l2, same_up_to_rounding


def _nhwc(a):
    return a.transpose(0, 2, 3, 1)


@pytest.mark.parametrize("name", SMALL)
def test_emulated_tiles_match_oracle(oracle32, name):
    from oracle.compose import geonet_loss
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(**CASES[name])
    e = emul.run_loss(inp, flags=3, want_maps=True)
    o = geonet_loss(oracle32, **oracle_inputs(inp), keep_maps=True)
    assert abs(e["loss_total"] - o["loss_total"]) < 1e-6 * abs(o["loss_total"])
    for s in range(inp.num_scales):
        assert not np.isnan(e["grad_depth"][s]).any()
        assert rel_l2(_nhwc(e["warped_fwd"][s]), o["fwd_warp"][s]) < 1e-5
        assert rel_l2(_nhwc(e["warped_bwd"][s]), o["bwd_warp"][s]) < 1e-5
        assert rel_l2(_nhwc(e["err_fwd"][s]), o["fwd_err"][s]) < 1e-5
        assert rel_l2(_nhwc(e["err_bwd"][s]), o["bwd_err"][s]) < 1e-5
        assert rel_l2(e["grad_depth"][s], o["grad_depth"][s]) < 1e-4
        np.testing.assert_allclose(e["smooth"][s], o["smooth"][s], rtol=1e-6)
    assert rel_l2(e["grad_poses"], o["grad_poses"]) < 1e-4
    assert rel_l2(e["grad_K_ms"], o["grad_K_ms"]) < 1e-4


def test_emulated_tiles_match_reference_golden_config1():
    """BASELINE config 1 against the reference's own numbers (tests/golden/config1_smooth.npz)."""
    from snde_b200.synthetic import make_inputs
    g = load_golden("config1_smooth")
    inp = make_inputs(**CASES["config1_smooth"])
    e = emul.run_loss(inp, flags=1)
    assert abs(e["loss_total"] - g["loss_total"]) < 1e-5 * abs(g["loss_total"])
    np.testing.assert_allclose(e["fwd_mean"], g["fwd_mean"], rtol=1e-5)
    np.testing.assert_allclose(e["bwd_mean"], g["bwd_mean"], rtol=1e-5)
    np.testing.assert_allclose(e["smooth"], g["smooth"], rtol=1e-5)
    for s in range(4):
        gd = e["grad_depth"][s].astype(np.float64)
        assert abs(np.sqrt((gd ** 2).sum()) - g[f"grad_depth_{s}_l2"]) < 1e-4 * g[f"grad_depth_{s}_l2"]
        rng = np.random.default_rng(1234 + s)
        # position-sensitive checksum; a few discontinuity pixels (see cases.close_except_flips) move it by ~1e-3
        assert abs((gd * rng.standard_normal(gd.shape)).sum() - g[f"grad_depth_{s}_proj"]) < 5e-3 * g[f"grad_depth_{s}_l2"]
    assert rel_l2(e["grad_poses"], g["grad_poses"]) < 1e-4     # measured 2.1e-5 (the reference's own distance from fp64: 1.5e-4)


def test_emulated_headline_config2_matches_reference_golden():
    """BASELINE config 2 at its full size (B=32) with the 64-row segments the library picks there (2,016 tasks) -- the exact
    decomposition bench.py times -- against the reference's own output at that size (tests/golden/config2_smooth.npz)."""
    from cases import FULL
    from snde_b200.synthetic import make_inputs
    g = load_golden("config2_smooth")
    inp = make_inputs(**FULL["config2_smooth"])
    e = emul.run_loss(inp, flags=1, seg_rows=64)
    assert abs(e["loss_total"] - g["loss_total"]) < 1e-5 * abs(g["loss_total"])
    np.testing.assert_allclose(e["fwd_mean"], g["fwd_mean"], rtol=1e-5)
    np.testing.assert_allclose(e["bwd_mean"], g["bwd_mean"], rtol=1e-5)
    np.testing.assert_allclose(e["smooth"], g["smooth"], rtol=1e-5)
    for s in range(4):
        gd = e["grad_depth"][s].astype(np.float64)
        assert not np.isnan(gd).any()
        assert abs(np.sqrt((gd ** 2).sum()) - g[f"grad_depth_{s}_l2"]) < 1e-4 * g[f"grad_depth_{s}_l2"]
    assert rel_l2(e["grad_poses"], g["grad_poses"]) < 1e-4     # measured 6.5e-5 (the reference's own distance from fp64: 2.1e-4)


def test_emulated_tiles_match_oracle_config1(oracle32):
    """config 1 shape, every depth-gradient pixel against the oracle (discontinuity pixels set aside)"""
    from oracle.compose import geonet_loss
    from snde_b200.synthetic import make_inputs
    from cases import close_except_flips
    inp = make_inputs(**CASES["config1_smooth"])
    e = emul.run_loss(inp, flags=1)
    o = geonet_loss(oracle32, **oracle_inputs(inp))
    assert abs(e["loss_total"] - o["loss_total"]) < 1e-6 * abs(o["loss_total"])
    for s in range(4):
        ok, rel, frac = close_except_flips(e["grad_depth"][s], o["grad_depth"][s])
        assert ok, (s, rel, frac)
    assert rel_l2(e["grad_poses"], o["grad_poses"]) < 1e-4


def test_emulated_masks_and_flags(oracle32):
    from oracle.compose import geonet_loss
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(**CASES["small_smooth"])
    rng = np.random.default_rng(5)
    masks = []
    for s in range(inp.num_scales):
        _, _, h, w = inp.tgt_pyr[s].shape
        masks.append((rng.random((inp.B * inp.S, h, w), np.float32), rng.random((inp.B * inp.S, h, w), np.float32)))
    e = emul.run_loss(inp, flags=1, masks=masks)
    o = geonet_loss(oracle32, **oracle_inputs(inp), masks=masks)
    assert abs(e["loss_total"] - o["loss_total"]) < 1e-6 * abs(o["loss_total"])
    for s in range(inp.num_scales):
        assert rel_l2(e["grad_depth"][s], o["grad_depth"][s]) < 1e-4
    assert rel_l2(e["grad_poses"], o["grad_poses"]) < 1e-4
    # photometric only / smoothness only add up to the total
    ep = emul.run_loss(inp, flags=1 | 8)
    es = emul.run_loss(inp, flags=1 | 4)
    et = emul.run_loss(inp, flags=1)
    assert abs(ep["loss_total"] + es["loss_total"] - et["loss_total"]) < 1e-6
    for s in range(inp.num_scales):
        assert rel_l2(ep["grad_depth"][s] + es["grad_depth"][s], et["grad_depth"][s]) < 1e-6
    # forward-only mode leaves the gradient buffers untouched and gives the same loss
    ef = emul.run_loss(inp, flags=0)
    assert abs(ef["loss_total"] - et["loss_total"]) < 1e-7
    assert np.isnan(ef["grad_depth"][0]).all()


def test_emulated_strided_views():
    """NHWC memory handed over as permuted views (the reference's layout) gives identical results"""
    import torch
    from snde_b200.synthetic import make_inputs, LossInputs
    inp = make_inputs(**CASES["small_noise_large_motion"])
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    inp2 = LossInputs([nhwc(t) for t in inp.tgt_pyr], [nhwc(t) for t in inp.src_pyr], inp.depth_pyr, inp.poses, inp.K, inp.K_ms)
    a, b = emul.run_loss(inp, flags=1), emul.run_loss(inp2, flags=1)
    assert a["loss_total"] == b["loss_total"]
    for s in range(inp.num_scales):
        assert np.array_equal(a["grad_depth"][s], b["grad_depth"][s])


@pytest.mark.parametrize("seg_rows", [5, 8, 13])
def test_emulated_segment_seams(seg_rows):
    """Cutting the strips into short row segments (seams every few rows) must not change a single depth-gradient bit:
    every inner pixel is still produced by exactly one lane from the same rolling sums."""
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(**CASES["small_noise_large_motion"])
    a, b = emul.run_loss(inp, flags=3), emul.run_loss(inp, flags=3, seg_rows=seg_rows)
    assert abs(a["loss_total"] - b["loss_total"]) < 1e-6 * abs(a["loss_total"])
    for s in range(inp.num_scales):
        assert np.array_equal(a["grad_depth"][s], b["grad_depth"][s])
    assert rel_l2(b["grad_poses"], a["grad_poses"]) < 1e-5
    assert rel_l2(b["grad_K_ms"], a["grad_K_ms"]) < 1e-5


def test_emulated_mixed_task_table():
    """long segments for most owners, short tail-filling segments for the rest: same bits"""
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(**CASES["small_smooth"])
    a = emul.run_loss(inp, flags=1)
    for fine_rows, pm in ((5, 500), (8, 1000), (4, 300)):
        b = emul.run_loss(inp, flags=1, seg_rows=16, fine_rows=fine_rows, fine_permille=pm)
        assert abs(a["loss_total"] - b["loss_total"]) < 1e-6 * abs(a["loss_total"])
        for s in range(inp.num_scales):
            assert np.array_equal(a["grad_depth"][s], b["grad_depth"][s])
        assert rel_l2(b["grad_poses"], a["grad_poses"]) < 1e-5


def test_emulated_variants_selected():
    """planar inputs take the FAST (64-bit pair access) variant, permuted NHWC views and odd widths the generic one"""
    import torch
    from snde_b200.synthetic import make_inputs, LossInputs
    inp = make_inputs(**CASES["small_smooth"])
    assert emul.run_loss(inp, flags=1)["variant"] == 4
    assert emul.run_loss(inp, flags=0)["variant"] == 1
    assert emul.run_loss(inp, flags=3)["variant"] == 7          # extras on planar inputs: general + FAST
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    inp2 = LossInputs([nhwc(t) for t in inp.tgt_pyr], [nhwc(t) for t in inp.src_pyr], inp.depth_pyr, inp.poses, inp.K, inp.K_ms)
    assert emul.run_loss(inp2, flags=1)["variant"] == 5


def test_emulated_odd_sizes(oracle32):
    """odd width / height (generic variant): columns and rows that do not fill a lane pair or a segment"""
    from oracle.compose import geonet_loss
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(B=1, S=2, H=37, W=67, num_scales=2, seed=11, smooth=True, motion="large")
    e = emul.run_loss(inp, flags=3, seg_rows=16)
    assert e["variant"] == 3                                    # odd width: generic strides
    o = geonet_loss(oracle32, **oracle_inputs(inp))
    assert abs(e["loss_total"] - o["loss_total"]) < 1e-6 * abs(o["loss_total"])
    for s in range(inp.num_scales):
        assert rel_l2(e["grad_depth"][s], o["grad_depth"][s]) < 1e-4
    assert rel_l2(e["grad_poses"], o["grad_poses"]) < 1e-4
    e1 = emul.run_loss(inp, flags=1, seg_rows=16)
    assert e1["variant"] == 5
    for s in range(inp.num_scales):
        assert same_up_to_rounding(e1["grad_depth"][s], e["grad_depth"][s])          # lean instance vs general instance


def test_emulated_disparity_input():
    """SNDE_LOSS_DISPARITY: depth = 1/disp inside the strip program, gradient w.r.t. the disparity = -(g * depth^2)"""
    import torch
    from snde_b200.synthetic import LossInputs, make_inputs
    inp = make_inputs(**CASES["small_smooth"])
    disp = [1.0 / d for d in inp.depth_pyr]
    depth = [1.0 / x for x in disp]                                  # what the reference forms (baseline.py:835)
    a = emul.run_loss(LossInputs(inp.tgt_pyr, inp.src_pyr, depth, inp.poses, inp.K, inp.K_ms), flags=1)
    b = emul.run_loss(LossInputs(inp.tgt_pyr, inp.src_pyr, disp, inp.poses, inp.K, inp.K_ms), flags=1 | 16)
    assert a["loss_total"] == b["loss_total"]
    for s in range(inp.num_scales):
        d = depth[s].numpy()
        want = (-a["grad_depth"][s]) * (d * d)
        assert np.array_equal(b["grad_depth"][s], want.astype(np.float32))
    assert np.array_equal(a["grad_poses"], b["grad_poses"])
    assert b["variant"] == 6
    c = emul.run_loss(LossInputs(inp.tgt_pyr, inp.src_pyr, disp, inp.poses, inp.K, inp.K_ms), flags=3 | 16)    # general instance, flag at run time
    assert c["variant"] == 7
    for s in range(inp.num_scales):
        assert same_up_to_rounding(c["grad_depth"][s], b["grad_depth"][s])          # general instance vs lean instance


@pytest.mark.parametrize("seed", range(8))
def test_emulated_random_shapes(oracle32, seed):
    """random small shapes (strip / segment / halo corner cases: w < 60, w just over a strip, h below a segment,
    S from 1 to 4, odd sizes, short segments) against the oracle"""
    from oracle.compose import geonet_loss
    from snde_b200.synthetic import make_inputs
    rng = np.random.default_rng(100 + seed)
    B, S = int(rng.integers(1, 3)), int(rng.integers(1, 5))
    ns = int(rng.integers(1, 3))
    H = int(rng.integers(6, 40)) * (2 ** (ns - 1)) // (2 ** (ns - 1)) + int(rng.integers(0, 2))
    W = int(rng.choice([8, 30, 58, 60, 61, 62, 64, 66, 118, 121, 124]))
    H = max(H, 4 * 2 ** (ns - 1)); W = max(W, 4 * 2 ** (ns - 1))
    inp = make_inputs(B=B, S=S, H=H, W=W, num_scales=ns, seed=seed, smooth=True, motion=("large" if seed % 2 else "small"))
    seg = int(rng.choice([0, 4, 7, 16]))
    e = emul.run_loss(inp, flags=int(rng.choice([1, 3])), seg_rows=seg, fine_rows=int(rng.choice([0, 3])), fine_permille=400)
    o = geonet_loss(oracle32, **oracle_inputs(inp))
    assert abs(e["loss_total"] - o["loss_total"]) < 2e-6 * abs(o["loss_total"]), (B, S, H, W, ns, seg)
    for s in range(inp.num_scales):
        assert not np.isnan(e["grad_depth"][s]).any()
        ok, rel, frac = close_except_flips(e["grad_depth"][s], o["grad_depth"][s], max_flip_frac=5e-3)
        assert ok, (s, rel, frac, (B, S, H, W, ns, seg))
    assert rel_l2(e["grad_poses"], o["grad_poses"]) < 2e-4


@pytest.mark.parametrize("name", SMALL)
def test_emulated_image_gradients_match_oracle(oracle32, name):
    """section 8(f3): the strip program's optional image-gradient outputs (dense part formed by the SSIM warp, scatter part by
    the geometry warp) against the oracle; the other outputs stay the same (up to the rounding
    difference between the general and the lean instance) when they are switched on"""
    from oracle.compose import geonet_loss
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(**CASES[name])
    e = emul.run_loss(inp, flags=1, want_image_grads=True)
    plain = emul.run_loss(inp, flags=1)
    o = geonet_loss(oracle32, **oracle_inputs(inp), want_grad_images=True)
    assert e["variant"] in (8, 9)
    assert abs(e["loss_total"] - plain["loss_total"]) < 1e-6 * abs(plain["loss_total"])        # (different order of the error sum)
    for s in range(inp.num_scales):
        assert same_up_to_rounding(e["grad_depth"][s], plain["grad_depth"][s])       # general (GIMG) instance vs lean instance
        for key in ("grad_tgt_img", "grad_src_img"):
            got = e[key][s].transpose(0, 2, 3, 1)
            assert not np.isnan(got).any()
            ok, rel, frac = close_except_flips(got, o[key][s], rtol=1e-4, max_flip_frac=2e-3)
            assert ok, (key, s, rel, frac)


@pytest.mark.parametrize("case", EDGE, ids=lambda c: f"B{c['B']}S{c['S']}_{c['H']}x{c['W']}x{c['num_scales']}")
def test_emulated_extreme_shapes_match_oracle(oracle32, case):
    """2x2 images, odd widths below one strip, 8 sources, one column past two strips, a few rows past a segment"""
    from oracle.compose import geonet_loss
    from snde_b200.synthetic import make_inputs
    inp = make_inputs(**case)
    o = geonet_loss(oracle32, **oracle_inputs(inp))
    for flags in (1, 3):
        e = emul.run_loss(inp, flags=flags)
        assert abs(e["loss_total"] - o["loss_total"]) < 1e-6 * abs(o["loss_total"])
        for s in range(inp.num_scales):
            assert not np.isnan(e["grad_depth"][s]).any()
            assert rel_l2(e["grad_depth"][s], o["grad_depth"][s]) < 1e-4
        assert rel_l2(e["grad_poses"], o["grad_poses"]) < 1e-4
        if flags & 2:
            assert rel_l2(e["grad_K_ms"], o["grad_K_ms"]) < 1e-4
