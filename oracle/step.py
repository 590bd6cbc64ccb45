"""Oracle restatement of the caller side of the hot path: ``masks.py`` column
masks, ``CSModel.set_input`` (model.py:89-121) and the ``reg='Rec'`` training
step ``forwardT`` + ``forwardR`` (model.py:142-169, 206-216).  Test infrastructure."""
import math
import random

import torch

from . import align, losses, varnet
from .signal import fft2, ifft2, rss


def equispaced_pruned(sparsity, shape, rng=random):
    """EquispacedMask masks.py:86-110 (DC at index 0; python ``random`` offset)."""
    center_len = round(shape * sparsity * 0.32)
    pruned = torch.zeros(shape, dtype=torch.bool)
    lo, hi = center_len // 2, center_len // 2 - center_len
    pruned[lo:hi] = True
    remaining = math.floor(sparsity * shape - center_len)
    interval = int((shape - center_len - 1) // (remaining - 1))
    start_max = (shape - center_len) - ((remaining - 1) * interval + 1)
    start = rng.randint(0, start_max)
    part = pruned[lo:hi].clone()
    part = torch.roll(part, part.shape[0] // 2)
    part[start:start + interval * remaining:interval] = False
    part = torch.roll(part, (part.shape[0] + 1) // 2)
    pruned[lo:hi] = part
    return pruned


def standard_pruned(sparsity, shape):
    """StandardMask masks.py:48-69 (uses the global torch RNG like the reference)."""
    center_len = round(shape * sparsity * 0.32)
    other = (sparsity * shape - center_len) / (shape - center_len)
    prob = torch.ones(shape) * 1.1
    prob[center_len // 2:center_len // 2 - center_len] = other
    thresh = torch.rand(shape)
    _, ind = torch.topk(prob - thresh, math.floor(sparsity * shape), dim=-1)
    return torch.ones(shape, dtype=torch.bool).scatter(-1, ind, torch.zeros(shape, dtype=torch.bool))


def set_input(img_full, img_aux, pruned):
    """CSModel.set_input model.py:108-121 -> dict of the ``img_*`` tensors the step uses."""
    k_full = fft2(img_full)
    k_sampled = k_full * (1 - pruned.to(k_full.real.dtype))      # multiply, model.py:113
    sampled = ifft2(k_sampled)
    return dict(img_full=img_full, img_aux=img_aux, img_k_full=k_full, img_k_sampled=k_sampled,
                img_sampled=sampled, img_full_rss=rss(img_full), img_sampled_rss=rss(sampled),
                img_aux_rss=rss(img_aux))


def registration_terms(full_rss, warped_rss, weight_lncc=0.0, weight_mi=0.0):
    """Optional similarity terms of BASELINE configs 3 / 5 between the target and the warped reference modality:
    lncc_loss (lnccloss.py:7-56) and ms_mi_loss (miloss.py:59-67).  The reference ships both losses but has them
    switched off in its live path (model.py:12); they enter ``loss_all`` with the given weights."""
    out, total = {}, 0.0
    if weight_lncc:
        out["loss_lncc"] = losses.lncc_loss(full_rss, warped_rss)
        total = total + out["loss_lncc"] * weight_lncc
    if weight_mi:
        out["loss_mi"] = losses.ms_mi_loss(full_rss, warped_rss)
        total = total + out["loss_mi"] * weight_mi
    return out, total


def rec_step(sd_T, sd_R, inp, pruned, shape, sparsity, num_cascades,
             weight_smooth=1000.0, weight_sim=1.0, training=True,
             sens_pools=4, pools=4, levels_T=4, weight_lncc=0.0, weight_mi=0.0):
    """forwardT + forwardR (model.py:142-169) under ``reg='Rec'``; returns dict with
    loss_all, loss_smooth, loss_sim, img_offset, img_grid, img_warped, img_rec."""
    offset, grid = align.spatial_transformer(sd_T, "", inp["img_aux"].abs(), inp["img_sampled"].abs(),
                                             training=training, num_levels=levels_T)
    warped = align.warp(inp["img_aux"].abs(), grid)
    loss_smooth = align.gradient_loss(offset)
    rec = varnet.varnet(sd_R, "", inp["img_k_sampled"], torch.logical_not(pruned), warped,
                        int(shape * sparsity * 0.32), num_cascades, sens_pools, pools, use_ref=True)
    loss_sim = losses.ssimloss(inp["img_full_rss"], rec)
    extra, extra_sum = registration_terms(inp["img_full_rss"], rss(warped), weight_lncc, weight_mi)
    loss_all = loss_smooth * weight_smooth + loss_sim * weight_sim + extra_sum
    return dict(loss_all=loss_all, loss_smooth=loss_smooth, loss_sim=loss_sim, img_offset=offset,
                img_grid=grid, img_warped=warped, img_rec=rec, **extra)
