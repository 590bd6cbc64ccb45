"""Diagnostic (GPU box): bisect backward error of the alignment network op by op.

Runs the golden ``align_s`` case twice through the SAME module code: once on the san_b200 kernels
(fp32) and once with every ``ops.X.apply`` swapped for a torch fp64 emulation; a tensor hook on every
op output records the gradient that reaches it.  Printing the per-op error in backward order shows the
first op whose *input* gradient is wrong.

Hooks the layer-by-layer ops: run with ``SAN_TC=0`` (the default fused tcgen05 path bypasses them)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_l2, sub  # noqa: E402
from spatialalignmentnetwork_b200 import ops, cross, unet as U, model as M  # noqa: E402

REC = []
DEV = "cuda" if torch.cuda.is_available() else "cpu"


def hooked(tag, fn):
    def wrapper(*a):
        out = fn(*a)
        idx = len(REC)
        REC.append([f"{idx:03d} {tag} {tuple(out.shape)}", out.detach(), None,
                    [t.detach().clone() if torch.is_tensor(t) else t for t in a] if tag == "BatchNormLReLU" else None])
        if out.requires_grad:
            out.register_hook(lambda g, i=idx: REC[i].__setitem__(2, g.detach().clone()))
        return out
    return wrapper


class Emu:
    Conv2d = staticmethod(lambda x, w, b: F.conv2d(x, w, b, padding=w.shape[-1] // 2))
    BatchNormLReLU = staticmethod(lambda y, g, b, rm, rv, tr, mom, eps, sl: F.leaky_relu(
        F.batch_norm(y, rm, rv, g, b, tr, mom, eps), sl))
    AvgPool2 = staticmethod(lambda x: F.avg_pool2d(x, 2, 2))
    Upsample2 = staticmethod(lambda x: x.repeat_interleave(2, 2).repeat_interleave(2, 3))
    add = staticmethod(lambda x, y: x + y)
    LReLU = staticmethod(lambda x, s: F.leaky_relu(x, s))


def run(emulate):
    REC.clear()
    g = load_golden("align_s")
    dt = torch.float64 if emulate else torch.float32
    st = cross.SpatialTransformer(1)
    st.load_state_dict(sub(g, "sd."))
    st = st.to(DEV).to(dt).train()
    saved = {}
    names = {"Conv2d": ops.Conv2d, "BatchNormLReLU": ops.BatchNormLReLU, "AvgPool2": ops.AvgPool2,
             "Upsample2": ops.Upsample2}
    for n, cls in names.items():
        saved[n] = cls.apply
        cls.apply = hooked(n, getattr(Emu, n) if emulate else saved[n])
    saved["add"] = ops.add
    ops.add = hooked("add", Emu.add if emulate else saved["add"])
    saved["lrelu"] = cross._LReLUFn.apply
    cross._LReLUFn.apply = hooked("lrelu", Emu.LReLU if emulate else saved["lrelu"])
    try:
        x = torch.cat([g["moving"], g["fixed"]], 1).to(DEV).to(dt)
        out = st.net(x)
        tgt = torch.linspace(-1, 1, out.numel(), device=DEV, dtype=dt).reshape(out.shape)
        ((out - tgt) ** 2).mean().backward()
    finally:
        for n, cls in names.items():
            cls.apply = saved[n]
        ops.add = saved["add"]
        cross._LReLUFn.apply = saved["lrelu"]
    return [list(r) for r in REC], {k: p.grad.detach().clone() for k, p in st.named_parameters()}


saved_bn = ops.BatchNormLReLU.apply


if __name__ == "__main__":
    ref, gr = run(True)
    ours, gp = run(False) if DEV == "cuda" else (ref, gr)
    assert len(ours) == len(ref)
    print("backward order: idx op shape | fwd err | err of grad wrt this op's OUTPUT")
    for a, b in reversed(list(zip(ours, ref))):
        ge = rel_l2(a[2], b[2]) if a[2] is not None and b[2] is not None else float("nan")
        print(f"{a[0]:45s} {rel_l2(a[1], b[1]):9.2e} {ge:9.2e}")

    print("\nreplay of every BatchNormLReLU on the fp64 run's own inputs: dy error ours / torch-cuda-fp32, cancellation")
    for r in reversed(ref):
        if r[3] is None or r[2] is None:
            continue
        y, gamma, beta, rm, rv, tr, mom, eps, sl = r[3]
        g = r[2]
        y64 = y.clone().requires_grad_(True)
        o64 = F.leaky_relu(F.batch_norm(y64, None, None, gamma, beta, True, mom, eps), sl)
        o64.backward(g)
        y32 = y.float().requires_grad_(True)
        o32 = F.leaky_relu(F.batch_norm(y32, None, None, gamma.float(), beta.float(), True, mom, eps), sl)
        o32.backward(g.float())
        yo = y.float().requires_grad_(True)
        oo = saved_bn(yo, gamma.float(), beta.float(), rm.float().clone(), rv.float().clone(), True, mom, eps, sl)
        oo.backward(g.float())
        gp_ = g * torch.where(o64 > 0, 1.0, sl)
        rstd = 1.0 / torch.sqrt(y.var((0, 2, 3), unbiased=False) + eps)
        canc = (gp_ * (gamma * rstd).view(1, -1, 1, 1)).norm() / y64.grad.norm()
        print(f"{r[0]:45s} ours {rel_l2(yo.grad, y64.grad):9.2e} torch32 {rel_l2(y32.grad, y64.grad):9.2e} "
              f"fwd ours {rel_l2(oo, o64):9.2e} cancellation {canc.item():8.1f}")
