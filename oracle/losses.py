"""Oracle restatement of reference ``ssimloss.py``, ``lnccloss.py``, ``miloss.py:6-67``
(test infrastructure)."""
import math

import torch
import torch.nn.functional as F


def _box(x, win, pad):
    w = torch.ones(1, 1, win, win, dtype=x.dtype, device=x.device)
    return F.conv2d(x, w, padding=pad)


def ssimloss(X, Y):
    """ssimloss.py:11-40: 7x7 uniform window, valid, k1=.01 k2=.03, range 1, cov_norm 49/48."""
    assert not torch.is_complex(X) and not torch.is_complex(Y)
    win, k1, k2 = 7, 0.01, 0.03
    NP = win ** 2
    cov_norm = NP / (NP - 1)
    C1, C2 = k1 ** 2, k2 ** 2
    w = torch.ones(1, 1, win, win).to(X) / NP
    ux, uy = F.conv2d(X, w), F.conv2d(Y, w)
    uxx, uyy, uxy = F.conv2d(X * X, w), F.conv2d(Y * Y, w), F.conv2d(X * Y, w)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    A1, A2 = 2 * ux * uy + C1, 2 * vxy + C2
    B1, B2 = ux ** 2 + uy ** 2 + C1, vx + vy + C2
    return 1 - ((A1 * A2) / (B1 * B2)).mean()


def lncc_loss(I, J, win=None):
    """lnccloss.py:7-56: 9x9 box sums, zero pad 4."""
    assert I.dim() == 4
    win = [9, 9] if win is None else win
    pad = win[0] // 2
    filt = torch.ones(1, 1, *win).to(I)
    conv = lambda t: F.conv2d(t, filt, stride=(1, 1), padding=(pad, pad))
    I_sum, J_sum = conv(I), conv(J)
    I2_sum, J2_sum, IJ_sum = conv(I * I), conv(J * J), conv(I * J)
    ws = float(win[0] * win[1])
    u_I, u_J = I_sum / ws, J_sum / ws
    cross = IJ_sum - u_J * I_sum - u_I * J_sum + u_I * u_J * ws
    I_var = I2_sum - 2 * u_I * I_sum + u_I * u_I * ws
    J_var = J2_sum - 2 * u_J * J_sum + u_J * u_J * ws
    cc = cross * cross / (I_var * J_var + 1e-5)
    return -1 * torch.mean(cc)


def gaussian_kernel_1d(sigma):
    """miloss.py:6-11."""
    ks = int(2 * math.ceil(sigma * 2) + 1)
    x = torch.linspace(-(ks - 1) // 2, (ks - 1) // 2, ks)
    k = 1.0 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return k / torch.sum(k)


def gaussian_smooth(img, sigma):
    """miloss.py:13-24."""
    k1 = gaussian_kernel_1d(sigma)
    k = torch.tensordot(k1, k1, 0)
    k = (k / torch.sum(k))[None, None].to(img)
    return F.conv2d(img, k, padding=k.shape[-1] // 2)


def _pyr_down(x, sigma):
    return F.avg_pool2d(gaussian_smooth(x, sigma), kernel_size=2, stride=2)


def ms_lncc_loss(I, J, win=None, ms=3, sigma=3):
    """lnccloss.py:58-65."""
    loss = lncc_loss(I, J, win)
    for _ in range(ms - 1):
        I, J = _pyr_down(I, sigma), _pyr_down(J, sigma)
        loss = loss + lncc_loss(I, J, win)
    return loss / ms


def _marginal(values, bins, sigma):
    """miloss.py:26-32."""
    norm1 = math.sqrt(2.0 * math.pi) * sigma
    p = torch.exp(-((values - bins).pow(2).div(2 * sigma ** 2))).div(norm1)
    p_n = p.mean(dim=1)
    p_n = p_n / (torch.sum(p_n) + 1e-10)
    return -(p_n * torch.log(p_n + 1e-10)).sum(), p


def _mi_one(I, J, bins, sigma):
    """miloss.py:36-46."""
    ent_I, p_I = _marginal(I.reshape(-1), bins, sigma)
    ent_J, p_J = _marginal(J.reshape(-1), bins, sigma)
    p_joint = torch.mm(p_I, p_J.transpose(0, 1)).div(2.0 * math.pi * sigma ** 2)
    p_joint = p_joint / (torch.sum(p_joint) + 1e-10)
    ent_joint = -(p_joint * torch.log(p_joint + 1e-10)).sum()
    return -(ent_I + ent_J - ent_joint)


def mi_loss(I, J, bins=64, sigma=1.0 / 64, minVal=0, maxVal=1):
    """miloss.py:49-57 (python loop over the batch)."""
    b = torch.linspace(minVal, maxVal, bins).to(I).unsqueeze(1)
    vals = [_mi_one(i, j, b, sigma) for i, j in zip(I, J)]
    return sum(vals) / len(vals)


def ms_mi_loss(I, J, bins=64, sigma=1.0 / 64, ms=3, smooth=3, minVal=0, maxVal=1):
    """miloss.py:59-67."""
    loss = mi_loss(I, J, bins, sigma, minVal, maxVal)
    for _ in range(ms - 1):
        I, J = _pyr_down(I, smooth), _pyr_down(J, smooth)
        loss = loss + mi_loss(I, J, bins, sigma, minVal, maxVal)
    return loss / ms
