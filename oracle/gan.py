"""Oracle restatement of reference ``gan.py:7-139`` (``NetG`` / ``NetD`` / ``loss_gan``), the spectral-norm
weight preprocessing it relies on (``torch.nn.utils.spectral_norm``: one power iteration per training
forward, eps 1e-12), the GAN parts of ``CSModel`` (``model.py:123-140,171-190,217-239``) and the numpy metrics
of ``metrics.py:23-68``.  Test infrastructure (see ``oracle/__init__.py``); plain functions over flat
``state_dict``s, pinned by ``tests/golden/gan_s.npz`` / ``mixed_step.npz`` / ``metrics.npz``."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import align, losses, varnet
from .signal import rss
from .step import registration_terms

SN_EPS = 1e-12


def sn_weight(sd, p, training):
    """Spectrally normalised filter of ``p + 'conv.'`` (gan.py:24).  Updates ``weight_u`` / ``weight_v`` of
    ``sd`` in place in training mode like the module's forward pre-hook; u, v are constants for autograd."""
    w = sd[p + "conv.weight_orig"]
    wm = w.reshape(w.shape[0], -1)
    u, v = sd[p + "conv.weight_u"], sd[p + "conv.weight_v"]
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS)
            sd[p + "conv.weight_u"], sd[p + "conv.weight_v"] = u, v
    sigma = torch.dot(u.detach(), torch.mv(wm, v.detach()))
    return w / sigma


def _bn(sd, p, x, training):
    """BatchNorm2d(affine, momentum 0.1, eps 1e-5) of a pre-activation block; running buffers replaced in ``sd``."""
    rm, rv = sd[p + "running_mean"].clone(), sd[p + "running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[p + "weight"], sd[p + "bias"], training, 0.1, 1e-5)
    if training:
        sd[p + "running_mean"], sd[p + "running_var"] = rm, rv
        sd[p + "num_batches_tracked"] = sd[p + "num_batches_tracked"] + 1
    return y


def conv_block(sd, p, x, training, norm=True, stride=1):
    """``Conv`` / ``ConvDown`` gan.py:10-28,43-46: norm -> ReLU -> spectral_norm(conv)."""
    if norm:
        x = _bn(sd, p + "norm_layer.", x, training)
    x = F.relu(x)
    w = sn_weight(sd, p, training)
    if stride == 2:
        return F.conv2d(x, w, sd[p + "conv.bias"], stride=2)
    return F.conv2d(x, w, sd[p + "conv.bias"], padding=1)


def _res(sd, p, x, n, training):
    out = x
    for i in range(n):
        out = conv_block(sd, f"{p}subnet.{i}.", out, training)
    return x + out


def _level(sd, p, x, depth, max_depth, training):
    """One CatSequential of NetG.__init__ gan.py:79-97 -> cat([module(x), x])."""
    y = conv_block(sd, p + "0.", x, training, stride=2)
    y = _res(sd, p + "1.", y, 2, training)
    if depth < max_depth:
        y = _level(sd, p + "2.module.", y, depth + 1, max_depth, training)
        y = conv_block(sd, p + "3.", y, training)
        y = _res(sd, p + "4.", y, 1, training)
    y = y.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)      # Up(): nearest x2
    return torch.cat([y, x], dim=1)


def netG(sd, p, x, num_levels, training=True):
    """NetG.forward gan.py:72-111 (``p`` ends with ``unet.``); ``num_levels`` = len(layers) - 1."""
    x = conv_block(sd, p + "0.", x, training)
    x = _res(sd, p + "1.", x, 1, training)
    x = _level(sd, p + "2.module.", x, 1, num_levels, training)
    x = conv_block(sd, p + "3.", x, training)
    x = _res(sd, p + "4.", x, 1, training)
    return conv_block(sd, p + "5.", x, training)


def netD(sd, p, x, blocks, training=True):
    """NetD.forward gan.py:113-129 (``p`` ends with ``net.``); ``blocks`` = convs per block, e.g. (2, 2, 2)."""
    i = 0
    for b, n in enumerate(blocks):
        for _ in range(n):
            x = conv_block(sd, f"{p}{i}.", x, training, norm=False)
            i += 1
        if b + 1 < len(blocks):
            x = F.avg_pool2d(x, 2)
        i += 1                                   # the Down() slot (the last one holds the 1-channel head)
    return conv_block(sd, f"{p}{i - 1}.", x, training, norm=False)


def loss_gan(predict, real=True, D_loss=True):
    """gan.py:131-137."""
    if D_loss:
        return torch.clamp(-predict if real else predict, min=-1).mean()
    assert not real
    return (-predict).mean()


def mixed_step(sd_T, sd_R, sd_G, sd_D, inp, pruned, shape, sparsity, num_cascades, g_levels, d_blocks,
               weight_smooth=1000.0, weight_sim=1.0, weight_gan=0.1, weight_gan_sim=1.0, with_R=True,
               sens_pools=4, pools=4, levels_T=4, weight_lncc=0.0, weight_mi=0.0):
    """forwardT + forwardG (+ forwardR) + forwardD(False), then forwardD(True) (model.py:123-190,217-260).
    Returns the generator-side and the discriminator-side results."""
    aux_abs = inp["img_aux"].abs()
    offset, grid = align.spatial_transformer(sd_T, "", aux_abs, inp["img_sampled"].abs(), training=True,
                                             num_levels=levels_T)
    warped = align.warp(aux_abs, grid)
    loss_smooth = align.gradient_loss(offset)
    aux_TR, aux_RT = torch.chunk(inp["img_aux_rss"], 2, dim=0)
    T = netG(sd_G, "unet.", aux_RT, g_levels)
    R, RT = torch.chunk(align.warp(torch.cat((aux_TR, T)), grid), 2)
    TR = netG(sd_G, "unet.", R, g_levels)
    synth, aligned = torch.cat((R, T), 0), torch.cat((TR, RT), 0)
    loss_gan_sim = F.l1_loss(aligned, inp["img_full_rss"])
    extra, extra_sum = registration_terms(inp["img_full_rss"], rss(warped), weight_lncc, weight_mi)
    loss_all = loss_smooth * weight_smooth + loss_gan_sim * weight_gan_sim + extra_sum
    out = dict(**extra, img_offset=offset, img_warped=warped, img_synth=synth, img_aligned=aligned, loss_smooth=loss_smooth,
               loss_gan_sim=loss_gan_sim)
    if with_R:
        rec = varnet.varnet(sd_R, "", inp["img_k_sampled"], torch.logical_not(pruned), warped,
                            int(shape * sparsity * 0.32), num_cascades, sens_pools, pools, use_ref=True)
        out["img_rec"] = rec
        out["loss_sim"] = losses.ssimloss(inp["img_full_rss"], rec)
        loss_all = loss_all + out["loss_sim"] * weight_sim
    fake = torch.cat((aligned, torch.zeros_like(aligned)), 1)
    real = torch.cat((inp["img_full_rss"], torch.zeros_like(inp["img_full_rss"])), 1)
    out["loss_gan_G"] = loss_gan(netD(sd_D, "net.", fake, d_blocks), real=False, D_loss=False)
    out["loss_G"] = loss_all + out["loss_gan_G"] * weight_gan

    def d_side():
        lf = loss_gan(netD(sd_D, "net.", fake.detach(), d_blocks), real=False, D_loss=True)
        lr = loss_gan(netD(sd_D, "net.", real.detach(), d_blocks), real=True, D_loss=True)
        return dict(loss_gan_Dfake=lf, loss_gan_Dreal=lr, loss_D=(lf + lr) * weight_gan)

    out["d_side"] = d_side
    return out


# ---------------------------------------------------------------------------------------- metrics.py
def metric_sums(gt, pred):
    """-> MSE, MAE, NMSE, PSNR (metrics.py:23-38; PSNR = skimage compare_psnr with data_range 1 over the batch)."""
    gt, pred = gt.double().numpy(), pred.double().numpy()
    mse = float(np.mean((gt - pred) ** 2))
    return dict(mse=mse, mae=float(np.mean(np.abs(gt - pred))),
                nmse=float(np.linalg.norm(gt - pred) ** 2 / np.linalg.norm(gt) ** 2), psnr=10 * math.log10(1.0 / mse))


def metric_mi(gt, pred, bins=64, minVal=0, maxVal=1):
    """metrics.mi metrics.py:54-68, histogram restated without np.histogram2d: bin k covers
    [e_k, e_{k+1}) of edges linspace(min, max, bins + 1), the last bin is closed, outliers are dropped."""
    vals = []
    edges = np.linspace(minVal, maxVal, bins + 1)
    for x, y in zip(gt.numpy(), pred.numpy()):
        x, y = x.ravel().astype(np.float64), y.ravel().astype(np.float64)
        ok = (x >= minVal) & (x <= maxVal) & (y >= minVal) & (y <= maxVal)
        bx = np.minimum(np.searchsorted(edges, x[ok], side="right") - 1, bins - 1)
        by = np.minimum(np.searchsorted(edges, y[ok], side="right") - 1, bins - 1)
        Pxy = np.zeros((bins, bins))
        np.add.at(Pxy, (bx, by), 1.0)
        Pxy = Pxy / (Pxy.sum() + 1e-10)
        PxPy = Pxy.sum(1)[:, None] * Pxy.sum(0)[None, :]
        nz = Pxy > 0
        vals.append(float((Pxy[nz] * np.log(Pxy[nz])).sum() - (Pxy[nz] * np.log(PxPy[nz])).sum()))
    return float(np.mean(vals))
