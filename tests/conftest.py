import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Load tests/golden/<name>.npz -> dict of torch tensors (+ nested 'sd.' dicts)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = torch.from_numpy(v.copy()) if v.ndim > 0 else torch.tensor(v.item())
    return out


def sub(d, prefix):
    """Extract {'<prefix>name': t} -> {'name': t}."""
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def rel_l2(a, b, floor=0.0):
    """||a-b|| / max(||b||, floor).  ``floor`` guards tensors that are zero in
    exact arithmetic (e.g. the gradient of a conv bias that feeds BatchNorm)."""
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    if torch.is_complex(a) or torch.is_complex(b):
        a, b = torch.view_as_real(a.to(torch.complex128)), torch.view_as_real(b.to(torch.complex128))
    a, b = a.double(), b.double()
    den = max(b.norm().item(), floor)
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def grad_floor(grads, frac=1e-3):
    """Absolute floor for per-parameter gradient comparisons: ``frac`` x the largest
    gradient norm in the set."""
    return frac * max(v.double().norm().item() for v in grads.values())
