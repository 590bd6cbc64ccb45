"""Pin oracle/augment.py against outputs of the unmodified reference ``augment.augment``
(tests/golden/make_golden_augment.py).  CPU only."""
import numpy as np
import torch

from oracle import augment as oaug
from conftest import load_golden, rel_l2

CASES = ("c_bs", "c_rigid", "r_bs", "r_rigid")


def test_augment_matches_reference():
    g = load_golden("augment")
    for key in CASES:
        img, theta = g[key + ".img"], g[key + ".theta"]
        ctrl = g.get(key + ".ctrl")
        out, grid = oaug.augment(img, theta, ctrl)
        assert rel_l2(grid, g[key + ".grid"]) < 1e-6, key
        assert rel_l2(out, g[key + ".out"]) < 2e-5, key
    assert rel_l2(oaug.sample_reflect(g["far.img"], g["far.grid"]), g["far.out"]) < 2e-5


def test_random_draws_follow_the_reference_streams():
    """rigid_theta consumes np.random like augment.py:12-13 (n rotations, then n translations)."""
    g = load_golden("augment")
    np.random.seed(40)          # seed of the first fixture case
    assert rel_l2(oaug.rigid_theta(3), g["c_bs.theta"]) < 1e-12


def test_bicubic_and_reflection_restatements_against_the_library():
    torch.manual_seed(3)
    c = torch.randn(2, 2, 9, 9)
    ref = torch.nn.functional.interpolate(c, size=(37, 50), mode="bicubic", align_corners=False)
    assert rel_l2(oaug.bicubic_upsample(c, 37, 50), ref) < 1e-6
    img = torch.rand(2, 3, 17, 23)
    grid = torch.rand(2, 11, 13, 2) * 9 - 4.5
    ref = torch.nn.functional.grid_sample(img, grid, mode="bilinear", padding_mode="reflection", align_corners=False)
    assert rel_l2(oaug.sample_reflect(img, grid), ref) < 1e-5


def _fake_call(name, *a):
    """The two augmentation kernels' documented semantics (include/san_b200.h) via the oracle, for host-logic tests."""
    if name == "augment_grid":
        theta, ctrl, G, grid, N, H, W = a
        g = oaug.affine_grid(theta.double(), H, W)
        if ctrl is not None:
            g = g + oaug.bicubic_upsample(ctrl.double(), H, W).permute(0, 2, 3, 1)
        grid.copy_(g.float())
    elif name == "warp_reflect":
        img, grid, out, N, C, H, W, Ho, Wo, k = a
        if k == 2:
            out.copy_(torch.view_as_real(torch.complex(oaug.sample_reflect(img[..., 0], grid),
                                                       oaug.sample_reflect(img[..., 1], grid))))
        else:
            out.copy_(oaug.sample_reflect(img, grid))
    else:
        raise KeyError(name)


def test_augment_module_host_logic(monkeypatch):
    """spatialalignmentnetwork_b200.augment (random draws, grid composition, complex handling, the --aux_aug modes,
    eval.py's augment_aux) with the two kernels replaced by their documented semantics: same seeds -> the
    reference's own outputs."""
    from spatialalignmentnetwork_b200 import augment as A
    monkeypatch.setattr(A, "call", _fake_call)
    g = load_golden("augment")
    for key, seed, bs in (("c_bs", 40, True), ("c_rigid", 45, False), ("r_bs", 49, True), ("r_rigid", 54, False)):
        np.random.seed(seed); torch.manual_seed(seed)         # seeds of tests/golden/make_golden_augment.py
        out, grid = A.augment(g[key + ".img"], rigid=True, bspline=bs)
        assert rel_l2(grid, g[key + ".grid"]) < 1e-6 and rel_l2(out, g[key + ".out"]) < 2e-5, key
    out, _ = A.augment(g["far.img"], rigid=False, bspline=False, grid=g["far.grid"])
    assert rel_l2(out, g["far.out"]) < 2e-5
    # PBSpline: one grid for the whole batch list; augment_aux: factor 1 reproduces a plain augment of the aux image
    np.random.seed(40); torch.manual_seed(40)
    a, b = A.augment_funcs["PBSpline"]([g["c_bs.img"], g["c_bs.img"].conj()])
    assert rel_l2(a, g["c_bs.out"]) < 2e-5 and rel_l2(b, g["c_bs.out"].conj()) < 2e-5
    np.random.seed(40); torch.manual_seed(40)
    full, aux = A.augment_aux((g["c_bs.img"] * 2, g["c_bs.img"]), factor=1)
    assert torch.equal(full, g["c_bs.img"] * 2) and rel_l2(aux, g["c_bs.out"]) < 5e-5
    np.random.seed(40); torch.manual_seed(40)
    _, aux3 = A.augment_aux((g["c_bs.img"], g["c_bs.img"]), factor=3)
    assert rel_l2(aux3, g["c_bs.out"]) > 1e-2                                      # a larger displacement
    assert A.center_crop(torch.zeros(1, 1, 352, 352), (320, 320)).shape[-2:] == (320, 320)
