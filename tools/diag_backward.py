"""Diagnostic (GPU box): where does backward error of the san_b200 modules come from?
Compares ours (CUDA fp32) and the CPU fp32 oracle against the CPU fp64 oracle, per tensor.
Not a test; run as ``python tools/diag_backward.py`` (``SAN_TC=0`` selects the layer-by-layer fp32 kernels)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_l2, sub  # noqa: E402
from oracle import varnet as ov, align as oa  # noqa: E402
from spatialalignmentnetwork_b200 import ops  # noqa: E402
from spatialalignmentnetwork_b200.varnet import Unet, NormUnet, VarNet  # noqa: E402
from spatialalignmentnetwork_b200.cross import SpatialTransformer  # noqa: E402


def report(title, ours, f32, f64, top=8):
    rows = []
    for k in f64:
        rows.append((rel_l2(ours[k], f64[k]), rel_l2(f32[k], f64[k]), k))
    rows.sort(reverse=True)
    print(f"== {title}: worst of {len(rows)} (ours-vs-fp64, cpu32-vs-fp64)")
    for a, b, k in rows[:top]:
        print(f"   {a:9.2e} {b:9.2e}  {k}")


def inorm_check():
    for shape, slope in (((2, 18, 64, 64), 0.2), ((2, 288, 4, 4), 0.2), ((2, 36, 160, 160), 0.2)):
        torch.manual_seed(1)
        x = torch.randn(*shape) * 2 + 0.7
        g = torch.randn(*shape)
        res = {}
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            xx = x.detach().clone().to(dt).requires_grad_(True)
            y = F.leaky_relu(F.instance_norm(xx, eps=1e-5), slope)
            (y * g.to(dt)).sum().backward()
            res[tag] = dict(y=y.detach(), dx=xx.grad)
        xc = x.detach().clone().cuda().requires_grad_(True)
        y = ops.InstanceNormLReLU.apply(xc, slope, 1e-5)
        (y * g.cuda()).sum().backward()
        res["ours"] = dict(y=y.detach(), dx=xc.grad)
        report(f"InstanceNormLReLU {shape}", res["ours"], res["f32"], res["f64"])


def unet_check(hw=64, chans=18, pools=4, N=2):
    torch.manual_seed(2)
    net = Unet(3, 2, chans, pools)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.randn(N, 3, hw, hw)
    g = torch.randn(N, 2, hw, hw)
    res = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        s = {k: v.clone().to(dt).requires_grad_(True) for k, v in sd.items()}
        xx = x.detach().clone().to(dt).requires_grad_(True)
        y = ov.unet(s, "", xx, pools)
        (y * g.to(dt)).sum().backward()
        res[tag] = {"y": y.detach(), "dx": xx.grad, **{"g." + k: v.grad for k, v in s.items()}}
    net.cuda()
    xc = x.detach().clone().cuda().requires_grad_(True)
    y = net(xc)
    (y * g.cuda()).sum().backward()
    res["ours"] = {"y": y.detach(), "dx": xc.grad, **{"g." + k: v.grad for k, v in net.named_parameters()}}
    report(f"Unet chans={chans} pools={pools} {hw}x{hw}", res["ours"], res["f32"], res["f64"])


def varnet_check(tag):
    g = load_golden(tag)
    nc, ch, pools, sch, sp = [int(v) for v in g["cfg"]]
    res = {}
    nlf = int(g["nlf"])
    for name, dt, cdt in (("f64", torch.float64, torch.complex128), ("f32", torch.float32, torch.complex64)):
        sd = {k: v.clone().to(dt).requires_grad_(True) for k, v in sub(g, "sd.").items()}
        ks = g["kspace"].clone().to(cdt).requires_grad_(True)
        ref = g["ref"].clone().to(dt).requires_grad_(True)
        rec = ov.varnet(sd, "", ks, ~g["pruned"], ref, nlf, nc, sp, pools, use_ref=True)
        ((rec - g["tgt"].to(dt)) ** 2).mean().backward()
        res[name] = {"rec": rec.detach(), "g_kspace": ks.grad, "g_ref": ref.grad,
                     **{"g." + k: v.grad for k, v in sd.items()}}
    net = VarNet(num_cascades=nc, sens_chans=sch, sens_pools=sp, chans=ch, pools=pools, use_ref=True)
    net.load_state_dict(sub(g, "sd."))
    net.cuda()
    ks = g["kspace"].cuda().requires_grad_(True)
    ref = g["ref"].cuda().requires_grad_(True)
    rec = net(ks, (~g["pruned"]).cuda(), ref, nlf)
    ((rec - g["tgt"].cuda()) ** 2).mean().backward()
    res["ours"] = {"rec": rec.detach(), "g_kspace": ks.grad, "g_ref": ref.grad,
                   **{"g." + k: v.grad for k, v in net.named_parameters()}}
    report(f"VarNet golden {tag}", res["ours"], res["f32"], res["f64"], top=10)
    gold = {k: g[k] for k in res["f64"]}
    report(f"VarNet golden {tag} [reference dump as 'cpu32']", res["ours"], gold, res["f64"], top=4)


def align_check(N=2, H=32, W=48, golden=True):
    if golden:
        g = load_golden("align_s")
        sd0 = sub(g, "sd.")
        moving, fixed, img, tgt = g["moving"], g["fixed"], g["img"], g["tgt"]
    else:
        torch.manual_seed(3)
        st = SpatialTransformer(1)
        torch.nn.init.normal_(st.net[-1].weight, 0, 1e-2)
        sd0 = {k: v.detach().clone() for k, v in st.state_dict().items()}
        moving, fixed, img = torch.rand(N, 1, H, W), torch.rand(N, 1, H, W), torch.rand(N, 1, H, W)
        tgt = torch.rand(N, 1, H, W)
    res = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        sd = {k: (v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
        im = img.detach().clone().to(dt).requires_grad_(True)
        offset, grid = oa.spatial_transformer(sd, "", moving.to(dt), fixed.to(dt), training=True)
        warped = oa.warp(im, grid)
        loss = ((warped - tgt.to(dt)) ** 2).mean() + 1000.0 * oa.gradient_loss(offset)
        loss.backward()
        res[name] = {"offset": offset.detach(), "g_img": im.grad,
                     **{"g." + k: v.grad for k, v in sd.items() if v.requires_grad}}
    st = SpatialTransformer(1)
    st.load_state_dict(sd0)
    st.cuda().train()
    im = img.detach().clone().cuda().requires_grad_(True)
    offset, grid = st(moving.cuda(), fixed.cuda())
    warped = st.warp(im, grid)
    from spatialalignmentnetwork_b200.model import gradient_loss
    loss = ((warped - tgt.cuda()) ** 2).mean() + 1000.0 * gradient_loss(offset)
    loss.backward()
    res["ours"] = {"offset": offset.detach(), "g_img": im.grad,
                   **{"g." + k: v.grad for k, v in st.named_parameters()}}
    print(f"== SpatialTransformer N={N} {H}x{W} golden={golden}: all weights in module order")
    for k in res["f64"]:
        if k.endswith("bias") and ".0.bias" in k:
            continue   # conv biases feeding BatchNorm: exact gradient is zero
        print(f"   {rel_l2(res['ours'][k], res['f64'][k]):9.2e} {rel_l2(res['f32'][k], res['f64'][k]):9.2e}  {k} {tuple(res['f64'][k].shape)}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "align":
        align_check()
        align_check(N=4, H=64, W=64, golden=False)
        sys.exit(0)
    inorm_check()
    unet_check()
    unet_check(hw=160, N=1)
    varnet_check("varnet_s")
    varnet_check("varnet_p")
    align_check()
    align_check(N=4, H=64, W=64, golden=False)
