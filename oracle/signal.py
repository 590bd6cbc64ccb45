"""Oracle restatement of reference ``signal_utils.py`` (test infrastructure)."""
import torch


def fft2(x):
    """signal_utils.py:4-7 — orthonormal, unshifted 2-D FFT over the last two dims."""
    assert x.dim() == 4
    return torch.fft.fft2(x, norm="ortho")


def ifft2(x):
    """signal_utils.py:9-12."""
    assert x.dim() == 4
    return torch.fft.ifft2(x, norm="ortho")


def fftshift2(x):
    """signal_utils.py:14-17 — roll by floor(n/2)."""
    assert x.dim() == 4
    return torch.roll(x, (x.shape[-2] // 2, x.shape[-1] // 2), dims=(-2, -1))


def ifftshift2(x):
    """signal_utils.py:19-22 — roll by ceil(n/2)."""
    assert x.dim() == 4
    return torch.roll(x, ((x.shape[-2] + 1) // 2, (x.shape[-1] + 1) // 2), dims=(-2, -1))


def rss(x):
    """signal_utils.py:24-26 — L2 norm over dim 1 (complex in -> real out)."""
    assert x.dim() == 4
    # the library call is part of the semantics: its backward defines the
    # sub-gradient 0 at |x| = 0 (an explicit sqrt(sum) would give NaN there).
    return torch.linalg.vector_norm(x, ord=2, dim=1, keepdim=True)


def dft2_direct(x, inverse=False):
    """Library-free cross-check: explicit DFT-matrix product (fp64 recommended)."""
    H, W = x.shape[-2:]
    sign = 2j if inverse else -2j
    cd = x.dtype
    rd = torch.float64 if cd == torch.complex128 else torch.float32
    kh = torch.arange(H, dtype=rd)
    kw = torch.arange(W, dtype=rd)
    FH = torch.exp((sign * torch.pi / H) * torch.outer(kh, kh).to(cd)) / (H ** 0.5)
    FW = torch.exp((sign * torch.pi / W) * torch.outer(kw, kw).to(cd)) / (W ** 0.5)
    return FH @ x @ FW
