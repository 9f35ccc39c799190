"""CPU checks of the C-ABI library: it loads, exports every symbol include/snde.h declares, its struct
layout matches the ctypes mirror, host helpers are bit-exact, and argument validation returns the documented
codes BEFORE any launch (no compute calls are made here -- there is no GPU in this suite)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cases import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def abi():
    from snde_b200 import _abi, build
    build.build()
    return _abi


def test_exports_every_declared_symbol(abi):
    header = open(os.path.join(ROOT, "include", "snde.h")).read()
    declared = set(re.findall(r"SNDE_API[^;(]*?\b(snde_\w+)\s*\(", header))
    assert len(declared) >= 17
    lib = C.CDLL(abi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(abi.SIGNATURES), declared ^ set(abi.SIGNATURES)


def test_version_and_layout(abi):
    lib = abi.lib()
    assert lib.snde_version() == 200
    assert lib.snde_sizeof_loss_args() == C.sizeof(abi.SndeLossArgs)
    assert b"NULL" in lib.snde_strerror(-1) and lib.snde_strerror(0) == b"ok"


def test_grid_table_matches_reference_bits(abi):
    k = load_golden("kats")
    for key in k.files:
        if key.startswith("grid_"):
            n = int(key.split("_")[1])
            out = np.empty(n, np.float32)
            assert abi.lib().snde_fill_grid_table(n, out.ctypes.data_as(abi.c_float_p)) == 0
            assert np.array_equal(out, k[key]), n
    assert abi.lib().snde_fill_grid_table(0, out.ctypes.data_as(abi.c_float_p)) == -2
    assert abi.lib().snde_fill_grid_table(4, None) == -1


def test_workspace_size_monotone(abi):
    lib = abi.lib()
    assert lib.snde_loss_workspace_bytes(0, 2, 4) == 0
    a, b = lib.snde_loss_workspace_bytes(4, 2, 4), lib.snde_loss_workspace_bytes(32, 2, 4)
    assert 0 < a < b and a % 16 == 0 and b % 16 == 0


def test_argument_validation_without_launch(abi):
    lib = abi.lib()
    args = abi.SndeLossArgs()
    assert lib.snde_loss_fwdbwd(None, None) == -1
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -1          # poses / loss_out NULL
    dummy = np.zeros(64, np.float32)
    args.poses = dummy.ctypes.data; args.loss_out = dummy.ctypes.data
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -2          # B = 0
    args.B, args.S, args.num_scales = 1, 1, 7
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -2          # too many scales
    args.num_scales = 1
    args.S = 9
    assert lib.snde_loss_fwdbwd(C.byref(args), None) in (-2, -4)    # more than SNDE_MAX_SOURCES views (or the workspace check first)
    args.S = 1
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -4          # workspace missing
    ws = np.zeros(4096, np.float64)
    args.workspace = (ws.ctypes.data + 15) & ~15; args.workspace_bytes = 16
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -4          # too small
    args.workspace_bytes = 4096 * 8 - 16
    args.alpha = 2.0
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -5          # alpha out of range
    args.alpha = 0.85
    args.scale[0].h, args.scale[0].w = 1, 8
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -2          # h < 2 (smoothness undefined)
    args.scale[0].h = 8
    assert lib.snde_loss_fwdbwd(C.byref(args), None) == -1          # image pointers NULL
    # reference-signature entry points
    z = abi.SndeMap(); zi = abi.SndeImage()
    assert lib.snde_rigid_flow_fwd(1, 4, 4, z, None, None, None, None, None, None) == -1
    assert lib.snde_smooth_loss_fwd(1, 1, 4, z, zi, None, None) == -1
    m = abi.SndeMap(dummy.ctypes.data, 16, 4, 1); im = abi.SndeImage(dummy.ctypes.data, 48, 16, 4, 1)
    assert lib.snde_smooth_loss_fwd(1, 1, 4, m, im, dummy.ctypes.data, None) == -2
    assert lib.snde_bilinear_sampler_fwd(1, 3, 4, 4, im, 4, 4, dummy.ctypes.data, 0, None, None, 7, dummy.ctypes.data, None, None) == -5
    assert lib.snde_image_similarity_fwd(0, 4, 4, 0.85, im, im, dummy.ctypes.data, None) == -2


def test_num_tiles(abi):
    """task table: class-F tasks (B owners per scale) + class-R tasks (S*B owners per scale); a task is one strip of 60
    inner columns x one segment of 64 rows (32 when that would leave fewer than two tasks per resident warp pair)"""
    lib = abi.lib()
    dummy = np.zeros(64, np.float32)
    args = abi.SndeLossArgs()
    args.B, args.S, args.num_scales, args.alpha = 32, 2, 4, 0.85
    for s, (h, w) in enumerate([(128, 416), (64, 208), (32, 104), (16, 52)]):
        sc = args.scale[s]
        sc.h, sc.w = h, w
        for f in ("tgt", "src"):
            getattr(sc, f).ptr = dummy.ctypes.data
        sc.depth.ptr = dummy.ctypes.data; sc.K = dummy.ctypes.data; sc.xtab = dummy.ctypes.data; sc.ytab = dummy.ctypes.data
    per_img = [7 * 2, 4 * 1, 2 * 1, 1 * 1]
    assert lib.snde_loss_num_tiles(C.byref(args)) == sum(per_img) * 32 * 3
    args.B = 4                                        # config 1: fewer tasks than resident pairs even at 32 rows -> 16-row segments
    assert lib.snde_loss_num_tiles(C.byref(args)) == (7 * 8 + 4 * 4 + 2 * 2 + 1 * 1) * 4 * 3


def test_no_fallback_when_library_missing(abi, monkeypatch):
    monkeypatch.setattr(abi, "_lib", None)
    monkeypatch.setattr(abi, "LIB_PATH", "/nonexistent/libsnde.so")
    with pytest.raises(abi.SndeLibraryError):
        abi.lib()


def test_cpu_tensors_are_rejected():
    import torch
    from snde_b200.loss_functions import image_similarity, smooth_loss
    x = torch.rand(1, 8, 8, 3)
    with pytest.raises(RuntimeError):
        image_similarity(0.85, x, x)
    with pytest.raises(RuntimeError):
        smooth_loss(torch.rand(1, 8, 8, 1), x)


def test_product_never_imports_oracle():
    """the package must not route through the CPU oracle (or anything under tests/): grep the sources"""
    import pathlib
    import re
    pkg = pathlib.Path(__file__).resolve().parents[1] / "self-supervised-normal-depth-estimation_b200"
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+(oracle|emul|tests)\b", txt, re.M), f
        assert "snde_oracle" not in txt and "host_emul/emul.cpp" not in txt.replace("tests/host_emul/emul.cpp", ""), f


def test_hot_kernel_instance_has_no_register_spills(abi):
    """The lean <GRAD, FAST> strip kernel sits at the 168-register limit of 3 CTAs x 128 threads per SM; a handful of
    spilled registers in either warp's loop costs ~20 % of the run time (profiles/README.md).  ptxas' report of the
    in-tree build must show none."""
    from snde_b200 import build
    log_path = os.path.join(os.path.dirname(abi.LIB_PATH), "build", os.path.basename(abi.LIB_PATH).replace(".so", "_") + "snde_fused.ptxas.log")
    if not os.path.exists(log_path):
        build.build(force=True)
    log = open(log_path).read()
    m = re.search(r"Function properties for _Z17loss_strip_kernelILb1ELb0ELb0ELb1ELb0ELb0EE\w+\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\n[^\n]*Used (\d+) registers", log)
    assert m, "lean kernel instance not found in the ptxas report"
    stack, st, ld, regs = map(int, m.groups())
    assert st == 0 and ld == 0, (st, ld)
    assert regs <= 168


def test_hot_kernel_instance_waits_for_its_gathers_only_at_the_blend(abi):
    """ptxas signals every global load of the geometry warp on one scoreboard; an instruction that waits on it waits for all
    24 tap gathers of the next row.  The only legitimate one is the blend at the top of the next row step -- a second
    one (the first use of a loop-invariant load, a read-modify-write) once cost 7-11 % of the run time each
    (profiles/README.md).  Static check of the in-tree build's SASS control bits."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not installed")
    from snde_b200 import build
    obj = os.path.join(os.path.dirname(abi.LIB_PATH), "build", os.path.basename(abi.LIB_PATH).replace(".so", "_") + "snde_fused.o")
    if not os.path.exists(obj):
        build.build(force=True)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    try:
        from sass_gather_waits import analyse
    finally:
        sys.path.pop(0)
    lo, hi, sbs, waits = analyse(obj, "_Z17loss_strip_kernelILb1ELb0ELb0ELb1ELb0ELb0EEvN4snde7DevLossE")
    assert sum(sbs.values()) == 24, sbs                       # the gathers are the only global loads of the loop
    assert len(waits) == 1, waits
    assert waits[0][0] - lo < 0x200, waits                    # ... and it sits at the top of the row step

