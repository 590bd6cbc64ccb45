"""Drop-in for the reference ``augment.py`` (augment.py:7-66): random rigid + b-spline misalignment of a batch,
applied on the device once per training step (train.py:207-212).  Random numbers are drawn exactly like the
reference (``np.random.uniform`` for the rotations / translations, ``torch.rand`` on the image's device for the
9x9 control displacements), so a seeded run consumes the same streams; the grid construction (affine_grid +
bicubic up-sampling + add) and the reflection-padded bilinear sampling run in two san_b200 kernels."""
import numpy as np
import torch

from ._lib import call

ROTATION = 2 * np.pi * 0.005
TRANSLATION = 0.05
BSPLINE_SCALE = 50
CTRL = 9


def rigid_theta(n):
    """[n, 2, 3] affine matrices T @ R of augment.py:7-33 (rotation about the image centre, equal x / y shift)."""
    r_s = np.random.uniform(-ROTATION, ROTATION, n)
    t_s = np.random.uniform(-TRANSLATION, TRANSLATION, n)
    theta = np.zeros((n, 2, 3))
    for i, (r, t) in enumerate(zip(r_s, t_s)):
        theta[i] = [[np.cos(r), -np.sin(r), t], [np.sin(r), np.cos(r), t]]
    return theta


def bspline_ctrl(img):
    """Control displacements in [-1/50, 1/50) on a 9x9 lattice (augment.py:41-44)."""
    dt = img.real.dtype if torch.is_complex(img) else img.dtype
    return (torch.rand(img.shape[0], 2, CTRL, CTRL, device=img.device, dtype=dt) - 0.5) * 2 / BSPLINE_SCALE


def _grid(theta, ctrl, N, H, W, device):
    theta = torch.as_tensor(theta, dtype=torch.float32).to(device, non_blocking=True).contiguous()
    grid = torch.empty(N, H, W, 2, dtype=torch.float32, device=device)
    call("augment_grid", theta, None if ctrl is None else ctrl.float().contiguous(), CTRL, grid, N, H, W)
    return grid


def rigid_grid(img):
    N, _, H, W = img.shape
    return _grid(rigid_theta(N), None, N, H, W, img.device)


def bspline_grid(img):
    N, _, H, W = img.shape
    return _grid(np.zeros((N, 2, 3)), bspline_ctrl(img), N, H, W, img.device)


def sample(img, grid):
    """grid_sample(bilinear, reflection, align_corners=False); complex input: real and imaginary parts."""
    assert img.dim() == 4 and grid.dim() == 4 and grid.shape[-1] == 2 and grid.shape[0] == img.shape[0]
    k = 2 if torch.is_complex(img) else 1
    img = (img.resolve_conj() if k == 2 else img).contiguous()      # (a lazy .conj() view cannot be viewed as real)
    assert img.dtype in (torch.float32, torch.complex64), img.dtype
    N, C, H, W = img.shape
    out = torch.empty(N, C, grid.shape[1], grid.shape[2], dtype=img.dtype, device=img.device)
    call("warp_reflect", torch.view_as_real(img) if k == 2 else img, grid.float().contiguous(),
         torch.view_as_real(out) if k == 2 else out, N, C, H, W, grid.shape[1], grid.shape[2], k)
    return out


def augment(img, rigid=True, bspline=True, grid=None):
    """-> (warped image, sampling grid [N,H,W,2]) (augment.py:46-62)."""
    if grid is None:
        assert rigid is True
        N, _, H, W = img.shape
        theta = rigid_theta(N)
        grid = _grid(theta, bspline_ctrl(img) if bspline else None, N, H, W, img.device)
    else:
        assert rigid is False
        assert bspline is False
    return sample(img, grid), grid


def augment_aux(batch, factor=1):
    """eval.py:15-27 of the reference: misalign the auxiliary modality by ``factor`` times a random rigid + b-spline
    displacement (robustness evaluation); the identity grid comes from the same grid kernel."""
    assert factor > 0
    img_full, img_aux = batch
    N, _, H, W = img_aux.shape
    _, grid = augment(img_aux, rigid=True, bspline=True)
    theta_id = np.tile(np.array([[[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]]), (N, 1, 1))
    identity = _grid(theta_id, None, N, H, W, img_aux.device)
    grid = identity + (grid - identity) * factor
    img_aux, _ = augment(img_aux, rigid=False, bspline=False, grid=grid)
    return (img_full, img_aux)


# the four ``--aux_aug`` modes of train.py:35-59
def augment_None(batch):
    return batch


def augment_Rigid(batch):
    return [augment(x, rigid=True, bspline=False)[0] for x in batch]


def augment_BSpline(batch):
    return [augment(x, rigid=True, bspline=True)[0] for x in batch]


def augment_PBSpline(batch):
    out, grid = [], None
    for x in batch:
        if grid is None:
            x, grid = augment(x, rigid=True, bspline=True)
        else:
            x, _ = augment(x, rigid=False, bspline=False, grid=grid)
        out.append(x)
    return out


augment_funcs = {"None": augment_None, "Rigid": augment_Rigid, "BSpline": augment_BSpline, "PBSpline": augment_PBSpline}


def center_crop(data, shape):
    """volumefolder.center_crop (volumefolder.py:9-16): a view."""
    assert 0 < shape[0] <= data.shape[-2] and 0 < shape[1] <= data.shape[-1]
    w0, h0 = (data.shape[-2] - shape[0]) // 2, (data.shape[-1] - shape[1]) // 2
    return data[..., w0:w0 + shape[0], h0:h0 + shape[1]]
