"""Drop-in for the reference ``signal_utils.py``: ``fft2`` / ``ifft2`` / ``fftshift2`` / ``ifftshift2`` /
``rss`` on 4-D ``[N, C, H, W]`` tensors (reference signal_utils.py:4-26), with the transforms and the
root-sum-of-squares running on the san_b200 CUDA kernels.  Orthonormal transforms over the last two dims,
DC at index [0, 0] (no implicit shift); non-4-D input raises ``AssertionError`` like the reference."""
import torch

from . import ops


def _nchw(x, who):
    assert x.dim() == 4, f"{who}: expected a 4-D [N, C, H, W] tensor, got {tuple(x.shape)}"
    return x


def _roll_hw(x, forward):
    # pure index permutation (not on the hot path): fftshift rolls by floor(n/2), ifftshift by ceil(n/2)
    h, w = x.shape[-2:]
    step = (h // 2, w // 2) if forward else ((h + 1) // 2, (w + 1) // 2)
    return torch.roll(x, step, dims=(-2, -1))


def fft2(x):
    return ops.Fft2.apply(_nchw(x, "fft2"), False)


def ifft2(x):
    return ops.Fft2.apply(_nchw(x, "ifft2"), True)


def fftshift2(x):
    return _roll_hw(_nchw(x, "fftshift2"), True)


def ifftshift2(x):
    return _roll_hw(_nchw(x, "ifftshift2"), False)


def rss(x):
    """L2 norm over the coil dimension (dim 1, kept); complex in -> real out."""
    return ops.Rss.apply(_nchw(x, "rss"))
