"""Host-side 1-D column sampling masks with the reference's conventions (DC at index 0;
reference masks.py:7-125).  320-element init code: not accelerated, but it must produce the
same ``pruned`` pattern as the reference because the pattern is an input of the hot path."""
import math
import random

import torch


class Mask(torch.nn.Module):
    """Holds ``weight`` (learnable scores, unused by the fixed masks) and the boolean
    ``pruned`` buffer (True = column NOT sampled), like reference masks.py:7-46."""

    def __init__(self, shape):
        super().__init__()
        self.shape = shape
        self.weight = torch.nn.Parameter(torch.ones(shape))
        self.register_buffer("pruned", torch.zeros(shape, dtype=torch.bool))

    def forward(self, image):
        keep = torch.logical_not(self.pruned).to(image.real.dtype if torch.is_complex(image) else image.dtype)
        return image * keep

    @property
    def sparsity(self):
        return 1.0 - self.pruned.float().mean().item()


def _center_slice(shape, sparsity):
    center_len = round(shape * sparsity * 0.32)
    return center_len, center_len // 2, center_len // 2 - center_len


class StandardMask(Mask):
    """Random columns + fully-sampled centre (reference masks.py:48-69); consumes the global
    torch RNG exactly like the reference (one ``torch.rand(shape)``)."""

    def __init__(self, sparsity, shape):
        super().__init__(shape)
        center_len, lo, hi = _center_slice(shape, sparsity)
        other = (sparsity * shape - center_len) / (shape - center_len)
        prob = torch.ones(shape) * 1.1
        prob[lo:hi] = other
        thresh = torch.rand(shape)
        _, ind = torch.topk(prob - thresh, math.floor(sparsity * shape), dim=-1)
        self.pruned = torch.ones(shape, dtype=torch.bool).scatter(-1, ind, torch.zeros(shape, dtype=torch.bool))


class EquispacedMask(Mask):
    """Equispaced columns with a random offset from python's ``random`` + fully-sampled centre
    (reference masks.py:86-110)."""

    def __init__(self, sparsity, shape):
        super().__init__(shape)
        center_len, lo, hi = _center_slice(shape, sparsity)
        pruned = torch.zeros(shape, dtype=torch.bool)
        pruned[lo:hi] = True
        remaining = math.floor(sparsity * shape - center_len)
        interval = int((shape - center_len - 1) // (remaining - 1))
        start_max = (shape - center_len) - ((remaining - 1) * interval + 1)
        start = random.randint(0, start_max)
        part = pruned[lo:hi].clone()
        part = torch.roll(part, part.shape[0] // 2)
        part[start:start + interval * remaining:interval] = False
        part = torch.roll(part, (part.shape[0] + 1) // 2)
        pruned[lo:hi] = part
        self.pruned = pruned


class LowpassMask(Mask):
    """Centre columns only (reference masks.py:112-125)."""

    def __init__(self, sparsity, shape):
        super().__init__(shape)
        center_len = math.floor(shape * sparsity)                  # reference masks.py:121
        pruned = torch.zeros(shape, dtype=torch.bool)
        pruned[center_len // 2:center_len // 2 - center_len] = True
        self.pruned = pruned


class _Unsupported(dict):
    """``masks[name]`` with an explicit error for the reference's learned masks (LOUPE / Taylor pruning,
    reference masks.py:141-244): they sit before the hot path and are not part of this build."""

    def __missing__(self, name):
        raise KeyError(f"unsupported mask {name!r}: this build provides {sorted(self)} "
                       "(the learned 'loupe' / 'taylor' masks of the reference are out of scope)")


# 'mask' = the plain fully-sampled ``Mask(shape)`` the reference registers under that name (model.py:30-37)
masks = _Unsupported({"mask": Mask, "standard": StandardMask, "equispaced": EquispacedMask, "lowpass": LowpassMask})
