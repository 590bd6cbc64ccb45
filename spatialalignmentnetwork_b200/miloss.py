"""Drop-in for the reference ``miloss.py`` (miloss.py:6-67): Parzen-window mutual information.

The O(bins^2 * pixels) work — the Gaussian soft histograms and their joint product, forward
and backward — runs in the san_b200 kernels, batched over images (the reference loops over the
batch in Python, miloss.py:56); the entropy arithmetic on the resulting [N,64,64] / [N,64]
tables follows the reference formulas with torch ops.
"""
import math

import numpy as np
import torch

from . import ops


def gaussian_kernel_1d(sigma):
    kernel_size = int(2 * math.ceil(sigma * 2) + 1)
    x = torch.linspace(-(kernel_size - 1) // 2, (kernel_size - 1) // 2, kernel_size)
    kernel = 1.0 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return kernel / torch.sum(kernel)


def gaussian_kernel_2d(sigma):
    y_1 = gaussian_kernel_1d(sigma)
    y_2 = gaussian_kernel_1d(sigma)
    kernel = torch.tensordot(y_1, y_2, 0)
    return kernel / torch.sum(kernel)


def gaussian_smooth(img, sigma):
    kernel = gaussian_kernel_2d(sigma).to(img)
    assert img.shape[1] == 1, "gaussian_smooth: single-channel images (as the reference's conv2d requires)"
    return ops.Filter2d.apply(img, kernel.contiguous())


def _entropy(p):
    return -(p * torch.log(p + 1e-10)).sum(dim=tuple(range(1, p.dim())))


def mi_loss(I, J, bins=64, sigma=1.0 / 64, minVal=0, maxVal=1):
    N = I.shape[0]
    Iv, Jv = I.reshape(N, -1), J.reshape(N, -1)
    P = Iv.shape[1]
    joint, mI, mJ = ops.MiHist.apply(Iv, Jv, bins, float(sigma), float(minVal), float(maxVal))
    pI = mI / P                                            # p.mean(dim=1), miloss.py:30
    pI = pI / (pI.sum(dim=1, keepdim=True) + 1e-10)
    pJ = mJ / P
    pJ = pJ / (pJ.sum(dim=1, keepdim=True) + 1e-10)
    pj = joint / (2.0 * math.pi * sigma ** 2)              # miloss.py:42
    pj = pj / (pj.sum(dim=(1, 2), keepdim=True) + 1e-10)
    mi = -(_entropy(pI) + _entropy(pJ) - _entropy(pj))     # per image, miloss.py:46
    return mi.sum() / N                                     # miloss.py:57


def ms_mi_loss(I, J, bins=64, sigma=1.0 / 64, ms=3, smooth=3, minVal=0, maxVal=1):
    smooth_fn = lambda x: ops.AvgPool2.apply(gaussian_smooth(x, smooth))
    loss = mi_loss(I, J, bins=bins, sigma=sigma, minVal=minVal, maxVal=maxVal)
    for _ in range(ms - 1):
        I, J = map(smooth_fn, (I, J))
        loss = loss + mi_loss(I, J, bins=bins, sigma=sigma, minVal=minVal, maxVal=maxVal)
    return loss / ms
