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


def assert_close_to_fp64(ours, f32, f64, bar, factor=4.0, floor_frac=1e-3):
    """Noise-floor-aware comparison for gradients through deep normalised nets.

    ``ours`` / ``f32`` / ``f64``: dicts name -> tensor from the CUDA path, the CPU fp32 oracle and
    the CPU fp64 oracle.  fp32 evaluation of these nets is chaotic in places (LeakyReLU sign flips,
    BatchNorm / InstanceNorm over a dozen elements, parameters whose exact gradient is zero), so the
    bar for each tensor is ``max(bar, factor * error of the CPU fp32 oracle)`` measured against
    fp64, with an absolute floor of ``floor_frac`` x the largest fp64 gradient norm."""
    fl = floor_frac * max(v.double().norm().item() for v in f64.values())
    for name, ref in f64.items():
        e_ours = rel_l2(ours[name], ref, fl)
        e_f32 = rel_l2(f32[name], ref, fl)
        assert e_ours < max(bar, factor * e_f32), f"{name}: ours {e_ours:.2e} vs cpu-fp32 {e_f32:.2e} (bar {bar:.0e})"


def cosine(a, b):
    a = torch.as_tensor(a).detach().cpu().double().flatten()
    b = torch.as_tensor(b).detach().cpu().double().flatten()
    if torch.is_complex(a):
        a, b = torch.view_as_real(a).flatten(), torch.view_as_real(b.to(a.dtype)).flatten()
    den = (a.norm() * b.norm()).item()
    return (a @ b).item() / den if den > 0 else 1.0


def kink_tolerant_failures(ours, ref, bar, floor_frac=1e-3):
    """-> (global relative L2 error of all gradients concatenated, [(name, rel, cos, norm ratio)] of the tensors
    that fail the per-tensor criterion of ``assert_grads_kink_tolerant``)."""
    names = [k for k in ref if k in ours and ours[k] is not None]
    assert names, "no gradients to compare"
    def flat(t):
        t = torch.as_tensor(t).detach().cpu()
        return (torch.view_as_real(t.to(torch.complex128)) if torch.is_complex(t) else t.double()).flatten()

    cat = lambda d: torch.cat([flat(d[k]) for k in names])
    glob = rel_l2(cat(ours), cat(ref))
    fl = floor_frac * max(flat(ref[k]).norm().item() for k in names)
    bad = []
    for k in names:
        e = rel_l2(ours[k], ref[k], fl)
        if e < 4 * bar:
            continue
        na, nb = flat(ours[k]).norm().item(), flat(ref[k]).norm().item()
        c = cosine(ours[k], ref[k])
        if not (c > 0.98 and 0.75 < na / max(nb, 1e-30) < 1.3333):
            bad.append((k, e, c, na / max(nb, 1e-30)))
    return glob, bad


def assert_grads_kink_tolerant(ours, ref, bar, what="", floor_frac=1e-3):
    """Gradient comparison for whole networks of (Instance|Batch)Norm + (Leaky)ReLU layers.

    fp32 evaluations of such nets differ by "kink flips" (a pre-activation within rounding of zero takes the other
    activation branch), each of which moves a gradient by O(1/sqrt(#activations of the layer)) -- percent level on
    the tiny golden nets and different for every change of summation order.  Robust criteria:
      * all gradients concatenated: relative L2 error < ``bar`` (dominated by the well-conditioned bulk);
      * every tensor: relative L2 error (floored at ``floor_frac`` of the largest gradient norm: tensors below the
        floor are sums of cancelling terms whose value is noise, they are covered by the global criterion)
        < 4*bar, or the same direction (cosine > 0.98) with a norm within 25 %.
    All failing tensors are reported together."""
    glob, bad = kink_tolerant_failures(ours, ref, bar, floor_frac)
    assert glob < bar, f"{what}: global gradient error {glob:.2e} >= {bar:.0e}"
    assert not bad, f"{what}: global {glob:.2e}; failing tensors (name, rel, cos, norm ratio): " + \
        "; ".join(f"{k}: {e:.2e}, {c:.4f}, {r:.3f}" for k, e, c, r in bad)
