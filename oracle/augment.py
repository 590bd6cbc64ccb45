"""Oracle restatement of reference ``augment.py:7-66`` with the arithmetic of the library calls it makes spelled
out: ``affine_grid`` (align_corners=False), ``interpolate(mode='bicubic', align_corners=False)`` (cubic
convolution, A = -0.75, border-clamped taps) and ``grid_sample(bilinear, padding_mode='reflection',
align_corners=False)``.  Test infrastructure; pinned by tests/golden/augment.npz (outputs of the reference's
``augment`` under fixed seeds)."""
import numpy as np
import torch


def rigid_theta(n, rotation=2 * np.pi * 0.005, translation=0.05):
    """augment.py:7-33 (consumes np.random exactly like the reference: n rotations, then n translations)."""
    r_s = np.random.uniform(-rotation, rotation, n)
    t_s = np.random.uniform(-translation, translation, n)
    th = np.zeros((n, 2, 3))
    for i, (r, t) in enumerate(zip(r_s, t_s)):
        R = np.array([[np.cos(r), -np.sin(r), 0], [np.sin(r), np.cos(r), 0], [0, 0, 1]])
        T = np.array([[1, 0, t], [0, 1, t], [0, 0, 1]])
        th[i] = (T @ R)[:-1]
    return torch.as_tensor(th)


def affine_grid(theta, H, W):
    """[N,2,3] -> [N,H,W,2]; pixel centres x_j = (2j+1)/W - 1."""
    xs = (2 * torch.arange(W, dtype=theta.dtype) + 1) / W - 1
    ys = (2 * torch.arange(H, dtype=theta.dtype) + 1) / H - 1
    base = torch.stack([xs[None, :].expand(H, W), ys[:, None].expand(H, W), torch.ones(H, W, dtype=theta.dtype)], -1)
    return torch.einsum("hwk,nck->nhwc", base, theta)


def _cubic(t):
    A = -0.75
    f1 = lambda x: ((A + 2) * x - (A + 3)) * x * x + 1          # |x| <= 1
    f2 = lambda x: ((A * x - 5 * A) * x + 8 * A) * x - 4 * A    # 1 < |x| < 2
    return torch.stack([f2(t + 1), f1(t), f1(1 - t), f2(2 - t)], -1)


def bicubic_upsample(x, H, W):
    """[N,C,h,w] -> [N,C,H,W], align_corners=False."""
    N, C, h, w = x.shape

    def axis(n_in, n_out):
        src = n_in / n_out * (torch.arange(n_out, dtype=x.dtype) + 0.5) - 0.5
        i0 = torch.floor(src)
        idx = (i0[:, None].long() + torch.arange(-1, 3)[None, :]).clamp(0, n_in - 1)
        return idx, _cubic(src - i0)

    iy, cy = axis(h, H)
    ix, cx = axis(w, W)
    rows = (x[:, :, iy, :] * cy[None, None, :, :, None]).sum(3)            # [N,C,H,w]
    return (rows[:, :, :, ix] * cx[None, None, None, :, :]).sum(4)        # [N,C,H,W]


def _reflect_clip(c, size):
    mn, span = -0.5, float(size)
    c = (c - mn).abs()
    extra = torch.fmod(c, span)
    flips = torch.floor(c / span).long()
    r = torch.where(flips % 2 == 0, extra + mn, span - extra + mn)
    return r.clamp(0, size - 1)


def sample_reflect(img, grid):
    """grid_sample(bilinear, reflection, align_corners=False) for real img [N,C,H,W]."""
    N, C, H, W = img.shape
    ix = _reflect_clip(((grid[..., 0] + 1) * W - 1) / 2, W)
    iy = _reflect_clip(((grid[..., 1] + 1) * H - 1) / 2, H)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx1, wy1 = ix - x0, iy - y0
    flat = img.reshape(N, C, H * W)
    out = torch.zeros(N, C, *grid.shape[1:3], dtype=img.dtype)
    for dy, wy in ((0, 1 - wy1), (1, wy1)):
        for dx, wx in ((0, 1 - wx1), (1, wx1)):
            xi, yi = (x0 + dx).long(), (y0 + dy).long()
            ok = (xi < W) & (yi < H)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).reshape(N, 1, -1).expand(N, C, -1)
            out = out + torch.gather(flat, 2, idx).reshape(out.shape) * (wx * wy * ok.to(img.dtype))[:, None]
    return out


def augment(img, theta, ctrl=None):
    """augment.py:46-62 with the random draws passed in: theta [N,2,3], ctrl [N,2,9,9] or None -> (img, grid)."""
    rdt = img.real.dtype if torch.is_complex(img) else img.dtype
    N, _, H, W = img.shape
    grid = affine_grid(theta.to(rdt), H, W)
    if ctrl is not None:
        grid = grid + bicubic_upsample(ctrl.to(rdt), H, W).permute(0, 2, 3, 1)
    if torch.is_complex(img):
        return torch.complex(sample_reflect(img.real, grid), sample_reflect(img.imag, grid)), grid
    return sample_reflect(img, grid), grid
