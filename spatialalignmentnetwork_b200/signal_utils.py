"""Drop-in for the reference ``signal_utils.py`` (fft2 / ifft2 / fftshift2 / ifftshift2 / rss
on 4-D tensors, reference signal_utils.py:4-26) running on the san_b200 CUDA kernels."""
import torch

from . import ops


def fft2(x):
    assert len(x.shape) == 4
    return ops.Fft2.apply(x, False)


def ifft2(x):
    assert len(x.shape) == 4
    return ops.Fft2.apply(x, True)


def fftshift2(x):
    # index permutation only (reference signal_utils.py:14-17); not on the hot path
    assert len(x.shape) == 4
    return torch.roll(x, (x.shape[-2] // 2, x.shape[-1] // 2), dims=(-2, -1))


def ifftshift2(x):
    assert len(x.shape) == 4
    return torch.roll(x, ((x.shape[-2] + 1) // 2, (x.shape[-1] + 1) // 2), dims=(-2, -1))


def rss(x):
    assert len(x.shape) == 4
    return ops.Rss.apply(x)
