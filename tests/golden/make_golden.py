"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Imports the reference's own modules, runs them on seeded CPU inputs (fp32, the
default the reference uses) and stores inputs, weights, outputs and gradients
as small .npz files.  The committed .npz files are what tests read; the GPU box
never sees /root/reference.
"""
import os
import random
import sys

import numpy as np
import torch

REF = os.environ.get("SAN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def npz(name, **kw):
    out = {}
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in out.items() if v.ndim > 0 and not k.startswith("sd")})


def sd_np(module, prefix):
    return {prefix + k: v for k, v in module.state_dict().items()}


def seed(s):
    torch.manual_seed(s); random.seed(s); np.random.seed(s)


def crand(*shape):
    return torch.complex(torch.rand(*shape), torch.rand(*shape))


def main():
    import signal_utils, varnet, cross, ssimloss, lnccloss, miloss, masks

    # ---- signal_utils (a1, a2) -------------------------------------------------
    seed(1)
    for tag, shp in (("a", (2, 3, 16, 20)), ("b", (1, 2, 20, 23)), ("c", (1, 1, 320, 320))):
        x = torch.complex(torch.randn(*shp), torch.randn(*shp))
        extra = {} if tag == "c" else dict(
            ifft2=signal_utils.ifft2(x), fftshift2=signal_utils.fftshift2(x),
            ifftshift2=signal_utils.ifftshift2(x), rss_real=signal_utils.rss(x.real))
        npz(f"signal_{tag}", x=x, fft2=signal_utils.fft2(x), rss=signal_utils.rss(x), **extra)

    # ---- masks ------------------------------------------------------------------
    seed(2)
    out = {}
    for shape, sp in ((320, 0.25), (320, 0.125), (368, 0.25), (64, 0.25)):
        random.seed(100 + shape)
        out[f"equi_{shape}_{sp}"] = masks.EquispacedMask(sp, shape).pruned
        torch.manual_seed(100 + shape)
        out[f"std_{shape}_{sp}"] = masks.StandardMask(sp, shape).pruned
    npz("masks", **out)

    # ---- VarNetBlock DC arithmetic alone (a3) with an identity regulariser ---------
    seed(3)
    class Ident(torch.nn.Module):
        def forward(self, x, ref):
            return x * (0.5 + 0.25j)
    blk = varnet.VarNetBlock(Ident())
    with torch.no_grad():
        blk.dc_weight.fill_(0.7)
    N, C, H, W = 2, 3, 16, 20
    k, k0, S = crand(N, C, H, W), crand(N, C, H, W), crand(N, C, H, W)
    m = torch.rand(W) > 0.5
    k_nan = k.clone(); k_nan[0, 0, 0, (~m).nonzero()[0, 0]] = complex(float("nan"), 0)
    npz("dc_block", k=k, k0=k0, S=S, mask=m, dc_weight=blk.dc_weight,
        reduce=blk.sens_reduce(k, S), expand=blk.sens_expand(k[:, :1], S),
        out=blk(k, k0, m, S, None))

    # ---- small VarNet, fwd + bwd (a3-a7) ----------------------------------------------
    for tag, (N, C, H, W, nc, ch, pools, sch, sp) in {
        "varnet_s": (2, 1, 32, 48, 2, 4, 2, 2, 2),      # no padding (mult of 16)
        "varnet_p": (1, 3, 40, 36, 1, 4, 2, 2, 2),      # multi-coil + pad-to-16 path
    }.items():
        seed(4)
        net = varnet.VarNet(num_cascades=nc, sens_chans=sch, sens_pools=sp, chans=ch, pools=pools, use_ref=True)
        with torch.no_grad():
            for i, c in enumerate(net.cascades):
                c.dc_weight.fill_(0.8 + 0.1 * i)
        random.seed(7)
        pruned = masks.EquispacedMask(0.25, W).pruned if W >= 40 else (torch.rand(W) > 0.4)
        full = crand(N, C, H, W)
        ksamp = signal_utils.fft2(full) * (1 - pruned.float())
        ref = torch.rand(N, C, H, W, requires_grad=True)
        ksamp.requires_grad_(True)
        nlf = max(2, int(W * 0.25 * 0.32))
        rec = net(ksamp, ~pruned, ref, nlf)
        tgt = torch.rand_like(rec)
        loss = ((rec - tgt) ** 2).mean()
        loss.backward()
        sens = net.sens_net(ksamp.detach(), nlf)
        grads = {"g." + k_: v.grad for k_, v in net.named_parameters()}
        npz(tag, kspace=ksamp, pruned=pruned, ref=ref, nlf=nlf, rec=rec, tgt=tgt, loss=loss, sens=sens,
            g_kspace=ksamp.grad, g_ref=ref.grad, cfg=np.array([nc, ch, pools, sch, sp]),
            **sd_np(net, "sd."), **grads)

    # ---- SpatialTransformer fwd + bwd, train and eval (a8, a9, a10) ---------------------
    seed(5)
    import model_shim  # noqa: F401  (installs the skimage stub so `model` imports)
    import model as refmodel
    st = cross.SpatialTransformer(1)
    with torch.no_grad():
        torch.nn.init.normal_(st.net[-1].weight, 0, 1e-2)
        torch.nn.init.normal_(st.net[-1].bias, 0, 1e-2)
        for mod in st.modules():       # non-trivial BN affine / running stats
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
                mod.running_mean.uniform_(-0.1, 0.1); mod.running_var.uniform_(0.5, 1.5)
    sd0 = {k_: v.clone() for k_, v in sd_np(st, "sd.").items()}
    N, H, W = 2, 32, 48
    moving, fixed = torch.rand(N, 1, H, W), torch.rand(N, 1, H, W)
    img = torch.rand(N, 1, H, W, requires_grad=True)
    st.train()
    offset, grid = st(moving, fixed)
    warped = st.warp(img, grid)
    tgt = torch.rand_like(warped)
    ls = refmodel.gradient_loss(offset)
    loss = ((warped - tgt) ** 2).mean() + 1000.0 * ls
    loss.backward()
    grads = {"g." + k_: v.grad for k_, v in st.named_parameters()}
    sd1 = {k_: v for k_, v in sd_np(st, "sd_after.").items() if "running" in k_ or "num_batches" in k_}
    st.eval()
    with torch.no_grad():
        offset_e, grid_e = st(moving, fixed)
    npz("align_s", moving=moving, fixed=fixed, img=img, offset=offset, grid=grid, warped=warped, tgt=tgt,
        loss=loss, loss_smooth=ls, g_img=img.grad, offset_eval=offset_e, **sd0, **sd1, **grads)

    # ---- warp alone with out-of-range grids ---------------------------------------------
    seed(6)
    img = torch.rand(2, 3, 17, 23, requires_grad=True)
    grid = (torch.rand(2, 17, 23, 2) * 2.6 - 1.3).requires_grad_(True)
    out = torch.nn.functional.grid_sample(img, grid, align_corners=False)
    w = torch.rand_like(out)
    (out * w).sum().backward()
    npz("warp", img=img, grid=grid, out=out, w=w, g_img=img.grad, g_grid=grid.grad)

    # ---- losses (a11-a13) -----------------------------------------------------------------
    seed(7)
    for tag, shp in (("s", (2, 1, 40, 36)), ("l", (1, 1, 320, 320))):
        X = torch.rand(*shp, requires_grad=True)
        Y = (X.detach() * 0.8 + 0.2 * torch.rand(*shp)).requires_grad_(True)
        res = {}
        for name, fn in (("ssim", ssimloss.ssimloss), ("lncc", lnccloss.lncc_loss),
                         ("mslncc", lnccloss.ms_lncc_loss), ("mi", miloss.mi_loss),
                         ("msmi", miloss.ms_mi_loss)):
            if tag == "l" and name in ("msmi", "mslncc"):
                continue
            X.grad = Y.grad = None
            v = fn(X, Y)
            v.backward()
            res[name] = v; res["gX_" + name] = X.grad.clone(); res["gY_" + name] = Y.grad.clone()
        if tag == "s":
            res["gauss"] = miloss.gaussian_smooth(X.detach(), 3)
        npz(f"losses_{tag}", X=X, Y=Y, **res)

    # ---- the Rec step end to end on the reference CSModel (a14 + all) ---------------------
    seed(8)
    from basemodel import Config
    shape = 32
    cfg = Config(sparsity=0.25, lr=1e-4, shape=shape, coils=1, reg="Rec", mask="equispaced",
                 weight_smooth=1000.0, weight_gan=0.1, weight_gan_sim=1.0, weight_sim=1.0, use_amp=False)
    random.seed(11)
    # CSModel hard-codes 8 cascades / 18 chans; shrink via monkeypatch of VarNet ctor args only
    orig = refmodel.VarNet
    refmodel.VarNet = lambda **kw: orig(**{**kw, "num_cascades": 2, "chans": 4, "pools": 2,
                                          "sens_chans": 2, "sens_pools": 2})
    net = refmodel.CSModel(cfg)
    refmodel.VarNet = orig
    with torch.no_grad():
        torch.nn.init.normal_(net.net_T.net[-1].weight, 0, 1e-2)
    sdT0 = {k_: v.clone() for k_, v in sd_np(net.net_T, "sdT.").items()}
    sdR0 = {k_: v.clone() for k_, v in sd_np(net.net_R, "sdR.").items()}
    full, aux = crand(2, 1, shape, shape), crand(2, 1, shape, shape)
    net.set_input(full, aux)
    k_sampled = net.img_k_sampled.clone()
    # replicate update() for reg='Rec' without the optimiser step so grads can be dumped
    net.loss_all = 0
    net.forwardT(); net.forwardR()
    net.optim_T.zero_grad(); net.optim_R.zero_grad()
    loss_all = net.loss_all
    loss_all.backward()
    gT = {"gT." + k_: v.grad for k_, v in net.net_T.named_parameters()}
    gR = {"gR." + k_: v.grad for k_, v in net.net_R.named_parameters()}
    npz("rec_step", full=full, aux=aux, pruned=net.net_mask.pruned, k_sampled=k_sampled,
        img_sampled=net.img_sampled, loss_all=loss_all, loss_smooth=net.loss_smooth, loss_sim=net.loss_sim,
        img_offset=net.img_offset, img_warped=net.img_warped, img_rec=net.img_rec,
        **sdT0, **sdR0, **gT, **gR)


if __name__ == "__main__":
    main()
