"""The misalignment augmentation kernels (augment.py:7-66 of the reference) through the C ABI against the
reference's own outputs (tests/golden/augment.npz) and the CPU oracle.  Needs a GPU."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key", ["c_bs", "c_rigid", "r_bs", "r_rigid"])
def test_grid_and_sampling_vs_reference(key):
    from spatialalignmentnetwork_b200 import augment as A
    g = load_golden("augment")
    img = g[key + ".img"].cuda()
    N, _, H, W = img.shape
    ctrl = g[key + ".ctrl"].cuda() if key + ".ctrl" in g else None
    grid = A._grid(g[key + ".theta"].numpy(), ctrl, N, H, W, img.device)
    assert rel_l2(grid, g[key + ".grid"]) < 2e-6
    out = A.sample(img, grid)
    assert out.dtype == img.dtype and rel_l2(out, g[key + ".out"]) < 2e-5
    # the reference's own grid -> only the sampling kernel
    assert rel_l2(A.sample(img, g[key + ".grid"].cuda()), g[key + ".out"]) < 1e-5


def test_far_outside_grid_reflects():
    from spatialalignmentnetwork_b200 import augment as A
    g = load_golden("augment")
    out, grid = A.augment(g["far.img"].cuda(), rigid=False, bspline=False, grid=g["far.grid"].cuda())
    assert rel_l2(out, g["far.out"]) < 1e-5 and grid.shape == g["far.grid"].shape


def test_augment_api_and_random_streams():
    """augment() draws like the reference: np.random for the rigid part (same seed -> same matrices as the
    fixture), torch.rand on the device for the control points; PBSpline shares one grid across the batch list."""
    from oracle import augment as oaug
    from spatialalignmentnetwork_b200 import augment as A
    g = load_golden("augment")
    img = g["r_rigid.img"].cuda()
    np.random.seed(54)      # seed of the r_rigid case (tests/golden/make_golden_augment.py)
    out, grid = A.augment(img, rigid=True, bspline=False)
    assert rel_l2(grid, g["r_rigid.grid"]) < 2e-6 and rel_l2(out, g["r_rigid.out"]) < 2e-5
    torch.manual_seed(1); np.random.seed(1)
    full = torch.complex(torch.rand(4, 1, 352, 352), torch.rand(4, 1, 352, 352)).cuda()
    aux = torch.complex(torch.rand(4, 1, 352, 352), torch.rand(4, 1, 352, 352)).cuda()
    a, b = A.augment_funcs["PBSpline"]([full, aux])
    np.random.seed(1)
    th = oaug.rigid_theta(4)
    # same rigid draw; the b-spline part moves every pixel by at most ~1/50 (+ bicubic overshoot) on top of it
    rigid_only = oaug.affine_grid(th.float(), 352, 352)
    torch.manual_seed(1); np.random.seed(1)
    _, grid = A.augment(full, rigid=True, bspline=True)
    dev = (grid.cpu() - rigid_only).abs().max().item()
    assert 1e-3 < dev < 0.04, dev
    assert a.shape == full.shape and b.shape == aux.shape
    assert A.center_crop(a, (320, 320)).shape[-2:] == (320, 320)
    assert A.augment_funcs["None"]([full, aux])[0] is full
    r = A.augment_funcs["Rigid"]([full])[0]
    assert r.shape == full.shape and torch.isfinite(torch.view_as_real(r)).all()
