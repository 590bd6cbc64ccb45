"""Golden fixtures of the GAN branch and the evaluation metrics, generated from
the UNMODIFIED reference like make_golden.py (build container only; needs /root/reference):
    python tests/golden/make_golden_gan.py
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, crand, npz, sd_np, seed  # noqa: E402,F401  (also puts the reference on sys.path)

G_LAYERS = (4, 8, 12, 8)
D_LAYERS = ([4] * 2, [8] * 2, [8] * 2)


def randomise_bn(net):
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
                mod.running_mean.uniform_(-0.1, 0.1); mod.running_var.uniform_(0.5, 1.5)
            if isinstance(mod, torch.nn.Conv2d) and mod.bias is not None:
                mod.bias.uniform_(-0.1, 0.1)


def main():
    import model_shim  # noqa: F401
    import gan
    import metrics
    import model as refmodel

    # ---- NetG / NetD alone: two training forwards (power iteration + running stats twice, as in forwardG),
    # backward, then an eval forward ---------------------------------------------------------------------
    seed(21)
    G = gan.NetG(1, 1, G_LAYERS)
    D = gan.NetD(2, D_LAYERS)
    randomise_bn(G); randomise_bn(D)
    sdG0 = {k: v.clone() for k, v in sd_np(G, "sdG.").items()}
    sdD0 = {k: v.clone() for k, v in sd_np(D, "sdD.").items()}
    x1 = torch.rand(2, 1, 32, 48, requires_grad=True)
    x2 = torch.rand(2, 1, 32, 48, requires_grad=True)
    G.train(); D.train()
    y1 = G(x1)
    y2 = G(x2)
    d1 = D(torch.cat([y1, torch.zeros_like(y1)], 1))
    tg = torch.rand_like(y2)
    l_g = gan.loss_gan(d1, real=False, D_loss=False)
    l_1 = torch.nn.functional.l1_loss(y2, tg)
    loss = l_1 + 0.1 * l_g
    loss.backward()
    gG = {"gG." + k: v.grad.clone() for k, v in G.named_parameters()}
    gD = {"gD." + k: v.grad.clone() for k, v in D.named_parameters()}
    # discriminator hinge terms on detached inputs
    D.zero_grad()
    xr = torch.rand(2, 1, 32, 48)
    lf = gan.loss_gan(D(torch.cat([y1.detach(), torch.zeros_like(y1)], 1)), real=False, D_loss=True)
    lr_ = gan.loss_gan(D(torch.cat([xr * 3 - 1, torch.zeros_like(xr)], 1)), real=True, D_loss=True)
    (lf + lr_).backward()
    gD2 = {"gD2." + k: v.grad.clone() for k, v in D.named_parameters()}
    sdG1 = {k: v.clone() for k, v in sd_np(G, "sdG_after.").items() if "running" in k or "num_batches" in k or "weight_u" in k or "weight_v" in k}
    sdD1 = {k: v.clone() for k, v in sd_np(D, "sdD_after.").items() if "weight_u" in k or "weight_v" in k}
    G.eval(); D.eval()
    with torch.no_grad():
        y_eval = G(x1)
        d_eval = D(torch.cat([y_eval, torch.zeros_like(y_eval)], 1))
    npz("gan_s", x1=x1, x2=x2, y1=y1, y2=y2, d1=d1, tg=tg, xr=xr, l_g=l_g, l_1=l_1, lf=lf, lr=lr_, g_x1=x1.grad, g_x2=x2.grad,
        y_eval=y_eval, d_eval=d_eval, **sdG0, **sdD0, **sdG1, **sdD1, **gG, **gD, **gD2)

    # ---- the Mixed step on the reference CSModel (model.py:217-239), shrunk networks -----------------------
    seed(22)
    from basemodel import Config
    shape = 32
    cfg = Config(sparsity=0.25, lr=1e-4, shape=shape, coils=1, reg="Mixed", mask="equispaced",
                 weight_smooth=1000.0, weight_gan=0.1, weight_gan_sim=1.0, weight_sim=1.0, use_amp=False)
    random.seed(12)
    oV, oG, oD = refmodel.VarNet, refmodel.NetG, refmodel.NetD
    refmodel.VarNet = lambda **kw: oV(**{**kw, "num_cascades": 2, "chans": 4, "pools": 2, "sens_chans": 2, "sens_pools": 2})
    refmodel.NetG = lambda **kw: oG(**{**kw, "layers": G_LAYERS})
    refmodel.NetD = lambda **kw: oD(**{**kw, "layers": D_LAYERS})
    net = refmodel.CSModel(cfg)
    refmodel.VarNet, refmodel.NetG, refmodel.NetD = oV, oG, oD
    with torch.no_grad():
        torch.nn.init.normal_(net.net_T.net[-1].weight, 0, 1e-2)
    sd0 = {}
    for tag, m in (("sdT.", net.net_T), ("sdR.", net.net_R), ("sdG.", net.net_G), ("sdD.", net.net_D)):
        sd0.update({k: v.clone() for k, v in sd_np(m, tag).items()})
    full, aux = crand(4, 1, shape, shape), crand(4, 1, shape, shape)
    net.set_input(full, aux)
    net.loss_all = 0
    net.forwardT(); net.forwardG(); net.forwardR(); net.forwardD(D_loss=False)
    for o in (net.optim_T, net.optim_G, net.optim_R, net.optim_D):
        o.zero_grad()
    loss_G = net.loss_all
    loss_G.backward()
    grads = {}
    for tag, m in (("gT.", net.net_T), ("gR.", net.net_R), ("gG.", net.net_G)):
        grads.update({tag + k: v.grad.clone() for k, v in m.named_parameters()})
    net.loss_all = 0
    net.forwardD(D_loss=True)
    net.optim_D.zero_grad()
    loss_D = net.loss_all
    loss_D.backward()
    grads.update({"gD." + k: v.grad.clone() for k, v in net.net_D.named_parameters()})
    npz("mixed_step", full=full, aux=aux, pruned=net.net_mask.pruned, loss_G=loss_G, loss_D=loss_D,
        loss_smooth=net.loss_smooth, loss_sim=net.loss_sim, loss_gan_sim=net.loss_gan_sim, loss_gan_G=net.loss_gan_G,
        loss_gan_Dfake=net.loss_gan_Dfake, loss_gan_Dreal=net.loss_gan_Dreal,
        img_synth=net.img_synth, img_aligned=net.img_aligned, img_rec=net.img_rec, img_warped=net.img_warped,
        g_layers=np.array(G_LAYERS), **sd0, **grads)

    # ---- metrics.py (numpy / scipy parts; skimage is absent here: PSNR / SSIM are pinned by their definitions,
    # PSNR = 10 log10(1 / MSE) for data_range 1, SSIM = 1 - ssimloss) ---------------------------
    seed(23)
    gt = torch.rand(3, 1, 40, 36)
    pred = (gt * 0.7 + 0.3 * torch.rand(3, 1, 40, 36))
    pred[0, 0, 0, :5] = torch.tensor([1.0, 0.0, 1.5, -0.2, 0.999999])      # edges of the histogram range, outliers
    gt[1, 0, 3, :3] = torch.tensor([1.0, 0.0, 0.5])
    npz("metrics", gt=gt, pred=pred, mse=metrics.mse(gt, pred), mae=metrics.mae(gt, pred), nmse=metrics.nmse(gt, pred),
        mi=metrics.mi(gt, pred), mi_each=np.array([metrics.mi(gt[i:i + 1], pred[i:i + 1]) for i in range(3)]),
        mi_16=metrics.mi(gt, pred, bins=16))


if __name__ == "__main__":
    main()
