"""Parity report of one ``reg='Rec'`` training step (reference model.py:142-169, 206-216) at ANY configuration,
including the headline ones (12 cascades, 320x320; 15 coils, 640x368): the step is evaluated by the CPU oracle in
fp32 (= the reference's own arithmetic) and in fp64 (calibrates the fp32 noise floor), and a candidate's outputs and
parameter gradients are compared with both.  Test infrastructure: imported by ``tests/`` and by ``bench.py``'s
``cpu_baseline`` leg only (it never imports the product package; the candidate arrives as plain tensors).

Also restates the two image metrics of the headline metric "PSNR/SSIM vs ref" (reference metrics.py:35-43:
skimage PSNR over the whole batch with data_range 1; mean per-image skimage SSIM, 7x7 uniform window, which equals
``1 - ssimloss`` - SURVEY.md 8c)."""
import math
import time

import torch

from . import losses, step as ostep


def psnr(gt, pred):
    """metrics.py:35-38: 10 log10(1 / mse) over the whole 4-D batch."""
    mse = (gt.double() - pred.double()).pow(2).mean().item()
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)


def ssim(gt, pred):
    """metrics.py:40-43 (skimage structural_similarity, data_range 1) == 1 - ssimloss for equal-size images."""
    return 1.0 - losses.ssimloss(gt.double(), pred.double()).item()


def _trainable(k, v):
    return v.is_floating_point() and "running" not in k and "weight_u" not in k and "weight_v" not in k


def oracle_rec_step(sd_T, sd_R, full, aux, pruned, shape, sparsity, cascades, dtype, **weights):
    """One Rec step by the oracle in ``dtype``; -> (outputs dict, grads dict 'T.<name>' / 'R.<name>', seconds)."""
    t0 = time.perf_counter()
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    sds = []
    for d in (sd_T, sd_R):
        d = {k: (v.detach().cpu().clone().to(dtype) if v.is_floating_point() else v.detach().cpu().clone()) for k, v in d.items()}
        for k, v in d.items():
            if _trainable(k, v):
                v.requires_grad_(True)
        sds.append(d)
    inp = ostep.set_input(full.cpu().to(cdt), aux.cpu().to(cdt), pruned.cpu())
    out = ostep.rec_step(sds[0], sds[1], inp, pruned.cpu(), shape, sparsity, cascades, **weights)
    out["loss_all"].backward()
    grads = {}
    for tag, d in (("T.", sds[0]), ("R.", sds[1])):
        for k, v in d.items():
            if v.requires_grad and v.grad is not None:
                grads[tag + k] = v.grad
    out["img_full_rss"] = inp["img_full_rss"]
    return out, grads, time.perf_counter() - t0


def _rel(a, b, floor=0.0):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / max(b.norm().item(), floor, 1e-300)).item()


def _cat(d, names):
    return torch.cat([d[k].detach().cpu().double().flatten() for k in names])


def compare(cand_out, cand_grads, f32, g32, f64, g64, worst=6):
    """-> JSON-able report.  ``cand_*``: the candidate's outputs (img_rec, img_warped, img_offset, loss_all) and
    gradients keyed like ``oracle_rec_step``'s.  Every error is a relative L2 distance; gradients are reported
    against fp64 for the candidate AND for the fp32 oracle (the noise floor of fp32 evaluation itself)."""
    rep = {"forward": {}, "grad": {}}
    for k in ("img_rec", "img_warped", "img_offset"):
        rep["forward"][k] = {"vs_fp32": _rel(cand_out[k], f32[k]), "vs_fp64": _rel(cand_out[k], f64[k]),
                             "fp32_vs_fp64": _rel(f32[k], f64[k])}
    la, l32, l64 = float(cand_out["loss_all"]), f32["loss_all"].item(), f64["loss_all"].item()
    rep["forward"]["loss_all"] = {"vs_fp32": abs(la - l32) / abs(l32), "vs_fp64": abs(la - l64) / abs(l64),
                                  "fp32_vs_fp64": abs(l32 - l64) / abs(l64)}
    names = [k for k in g64 if k in cand_grads and cand_grads[k] is not None]
    missing = [k for k in g64 if k not in names]
    big = max(g64[k].double().norm().item() for k in names)
    fl = 1e-3 * big          # tensors below the floor are sums of cancelling terms (e.g. the exactly-zero dc_weight grad)
    per = {k: (_rel(cand_grads[k], g64[k], fl), _rel(g32[k], g64[k], fl)) for k in names}
    c, r32, r64 = _cat(cand_grads, names), _cat(g32, names), _cat(g64, names)
    glob_c, glob_32 = ((c - r64).norm() / r64.norm()).item(), ((r32 - r64).norm() / r64.norm()).item()
    cos = (c @ r64 / (c.norm() * r64.norm())).item()
    cos32 = (r32 @ r64 / (r32.norm() * r64.norm())).item()
    order = sorted(names, key=lambda k: -per[k][0])
    rep["grad"] = {
        "tensors": len(names), "missing": missing,
        "all_concatenated": {"vs_fp64": glob_c, "fp32_vs_fp64": glob_32, "ratio_to_fp32_floor": glob_c / max(glob_32, 1e-30),
                             "vs_fp32": ((c - r32).norm() / r32.norm()).item(), "cosine_vs_fp64": cos,
                             "fp32_cosine_vs_fp64": cos32},
        "per_tensor_median": {"vs_fp64": sorted(v[0] for v in per.values())[len(per) // 2],
                              "fp32_vs_fp64": sorted(v[1] for v in per.values())[len(per) // 2]},
        "worst": [{"name": k, "vs_fp64": per[k][0], "fp32_vs_fp64": per[k][1]} for k in order[:worst]],
    }
    rec, ref_rec, gt = cand_out["img_rec"].detach().cpu(), f32["img_rec"].detach(), f32["img_full_rss"]
    rep["image_metrics"] = {
        "psnr_rec_vs_reference_rec_db": psnr(ref_rec, rec), "ssim_rec_vs_reference_rec": ssim(ref_rec, rec),
        "psnr_rec_vs_full": psnr(gt, rec), "psnr_reference_rec_vs_full": psnr(gt, ref_rec),
        "ssim_rec_vs_full": ssim(gt, rec), "ssim_reference_rec_vs_full": ssim(gt, ref_rec),
    }
    return rep


def rec_step_report(sd_T, sd_R, full, aux, pruned, shape, sparsity, cascades, cand_out, cand_grads, **weights):
    """Run the oracle in fp32 and fp64 and compare the candidate; adds the oracle wall times (the fp32 time on
    ``full.shape[0]`` slices doubles as a CPU-baseline sample)."""
    f32, g32, t32 = oracle_rec_step(sd_T, sd_R, full, aux, pruned, shape, sparsity, cascades, torch.float32, **weights)
    f64, g64, t64 = oracle_rec_step(sd_T, sd_R, full, aux, pruned, shape, sparsity, cascades, torch.float64, **weights)
    rep = compare(cand_out, cand_grads, f32, g32, f64, g64)
    rep["oracle_seconds"] = {"fp32": round(t32, 2), "fp64": round(t64, 2)}
    rep["config"] = {"slices": int(full.shape[0]), "coils": int(full.shape[1]), "shape": [int(full.shape[2]), int(full.shape[3])],
                     "cascades": int(cascades), "reg": "Rec"}
    return rep
