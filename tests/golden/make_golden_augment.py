"""Golden fixture of the misalignment augmentation from the UNMODIFIED reference
``augment.augment`` (build container only; needs /root/reference):
    python tests/golden/make_golden_augment.py
The reference draws its random parameters internally (np.random.uniform, torch.rand); the fixture stores the
same draws, reproduced with the same seeds in the same order, next to the reference's outputs."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import npz  # noqa: E402  (also puts the reference on sys.path)


def draws(n, seed, bspline):
    """The reference's random draws for a batch of n under `seed` (augment.py:12-13,41-42)."""
    np.random.seed(seed); torch.manual_seed(seed)
    r_s = np.random.uniform(-2 * np.pi * 0.005, 2 * np.pi * 0.005, n)
    t_s = np.random.uniform(-0.05, 0.05, n)
    theta = np.stack([np.array([[np.cos(r), -np.sin(r), t], [np.sin(r), np.cos(r), t]]) for r, t in zip(r_s, t_s)])
    ctrl = (torch.rand(n, 2, 9, 9) - 0.5) * 2 / 50 if bspline else None
    return theta, ctrl


def main():
    import model_shim  # noqa: F401
    import augment as raug
    out = {}
    torch.manual_seed(31)
    imgs = {"c": torch.complex(torch.rand(3, 2, 40, 52), torch.rand(3, 2, 40, 52)), "r": torch.rand(2, 1, 33, 47)}
    for tag, img in imgs.items():
        for bs in (True, False):
            seed = 40 + len(out)
            theta, ctrl = draws(img.shape[0], seed, bs)
            np.random.seed(seed); torch.manual_seed(seed)
            res, grid = raug.augment(img, rigid=True, bspline=bs)
            key = f"{tag}_{'bs' if bs else 'rigid'}"
            out.update({key + ".img": img, key + ".theta": theta, key + ".out": res, key + ".grid": grid})
            if bs:
                out[key + ".ctrl"] = ctrl
    # a given grid far outside [-1, 1]: several reflections
    img = imgs["r"]
    grid = torch.rand(2, 33, 47, 2) * 7 - 3.5
    res, _ = raug.augment(img, rigid=False, bspline=False, grid=grid)
    out.update({"far.img": img, "far.grid": grid, "far.out": res})
    npz("augment", **out)


if __name__ == "__main__":
    main()
