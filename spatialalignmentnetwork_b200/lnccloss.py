"""Drop-in for the reference ``lnccloss.py`` (lnccloss.py:7-65) on the san_b200 LNCC kernel."""
import torch

from . import ops
from .miloss import gaussian_smooth


def lncc_loss(I, J, win=None):
    ndims = len(list(I.size())) - 2
    assert ndims == 2, "images should be 2 dimensions. found: %d" % ndims
    if win is None:
        win = [9] * ndims
    assert list(win) == [9, 9], "san_b200 lncc kernel is specialised for the reference's 9x9 window"
    return ops.LnccLoss.apply(I, J)


def ms_lncc_loss(I, J, win=None, ms=3, sigma=3):
    smooth_fn = lambda x: ops.AvgPool2.apply(gaussian_smooth(x, sigma))
    loss = lncc_loss(I, J, win)
    for _ in range(ms - 1):
        I, J = map(smooth_fn, (I, J))
        loss = loss + lncc_loss(I, J, win)
    return loss / ms
