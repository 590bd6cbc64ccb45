"""Pin the oracle (oracle/) against golden vectors minted from the unmodified
reference (tests/golden/make_golden.py).  CPU only.

Tolerances: the oracle runs the same fp32 CPU library as the reference, so most
results are bit-identical; where the restatement orders operations differently
(explicit grid_sample, instance-norm) the bar is 2e-5 rel-L2, far below the
1e-3 parity bar the CUDA path is held to.
"""
import random

import pytest
import torch

import oracle
from oracle import align, losses, signal, step, varnet
from conftest import grad_floor, load_golden, rel_l2, sub

TOL = 2e-5


def test_signal_utils():
    for tag in "abc":
        g = load_golden(f"signal_{tag}")
        x = g["x"]
        assert rel_l2(signal.fft2(x), g["fft2"]) < 1e-6
        assert rel_l2(signal.rss(x), g["rss"]) < 1e-6
        if tag != "c":
            assert rel_l2(signal.ifft2(x), g["ifft2"]) < 1e-6
            assert torch.equal(signal.fftshift2(x), g["fftshift2"])
            assert torch.equal(signal.ifftshift2(x), g["ifftshift2"])
            assert rel_l2(signal.rss(x.real), g["rss_real"]) < 1e-6
            # library-free cross-check of the FFT definition (fp64 DFT matrices)
            xd = x.to(torch.complex128)
            assert rel_l2(signal.dft2_direct(xd), g["fft2"]) < 1e-6
            assert rel_l2(signal.dft2_direct(xd, inverse=True), g["ifft2"]) < 1e-6


def test_masks_bit_exact():
    g = load_golden("masks")
    for shape, sp in ((320, 0.25), (320, 0.125), (368, 0.25), (64, 0.25)):
        random.seed(100 + shape)
        assert torch.equal(step.equispaced_pruned(sp, shape), g[f"equi_{shape}_{sp}"])
        torch.manual_seed(100 + shape)
        assert torch.equal(step.standard_pruned(sp, shape), g[f"std_{shape}_{sp}"])


def test_dc_block():
    g = load_golden("dc_block")
    k, k0, S, m = g["k"], g["k0"], g["S"], g["mask"]
    red = (signal.ifft2(k) * S.conj()).sum(1, keepdim=True)
    assert rel_l2(red, g["reduce"]) < 1e-6
    assert rel_l2(signal.fft2(k[:, :1] * S), g["expand"]) < 1e-6
    x = red * (0.5 + 0.25j)
    out = k - torch.where(m, k - k0, torch.zeros(1, 1, 1, 1).to(k)) * g["dc_weight"] - signal.fft2(x * S)
    assert rel_l2(out, g["out"]) < 1e-6


@pytest.mark.parametrize("tag", ["varnet_s", "varnet_p"])
def test_varnet_fwd_bwd(tag):
    g = load_golden(tag)
    nc, ch, pools, sch, sp = [int(v) for v in g["cfg"]]
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sub(g, "sd.").items()}
    ks = g["kspace"].clone().requires_grad_(True)
    ref = g["ref"].clone().requires_grad_(True)
    nlf = int(g["nlf"])
    sens = varnet.sens_model(sd, "sens_net.", ks.detach(), nlf, sp)
    assert rel_l2(sens, g["sens"]) < TOL
    rec = varnet.varnet(sd, "", ks, ~g["pruned"], ref, nlf, nc, sp, pools, use_ref=True)
    assert rel_l2(rec, g["rec"]) < TOL
    loss = ((rec - g["tgt"]) ** 2).mean()
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) < 1e-5 * abs(g["loss"].item())
    assert rel_l2(ks.grad, g["g_kspace"]) < 1e-4
    assert rel_l2(ref.grad, g["g_ref"]) < 1e-4
    for name, gg in sub(g, "g.").items():
        assert rel_l2(sd[name].grad, gg) < 1e-4, name


def test_align_fwd_bwd():
    g = load_golden("align_s")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k)
          for k, v in sub(g, "sd.").items()}
    img = g["img"].clone().requires_grad_(True)
    offset, grid = align.spatial_transformer(sd, "", g["moving"], g["fixed"], training=True)
    assert rel_l2(offset, g["offset"]) < TOL
    assert rel_l2(grid, g["grid"]) < TOL
    warped = align.warp(img, grid)
    assert rel_l2(warped, g["warped"]) < TOL
    ls = align.gradient_loss(offset)
    assert abs(ls.item() - g["loss_smooth"].item()) < 1e-5 * abs(g["loss_smooth"].item())
    loss = ((warped - g["tgt"]) ** 2).mean() + 1000.0 * ls
    loss.backward()
    assert rel_l2(img.grad, g["g_img"]) < 1e-4
    fl = grad_floor(sub(g, "g."))
    for name, gg in sub(g, "g.").items():
        assert rel_l2(sd[name].grad, gg, fl) < 2e-4, name
    for name, v in sub(g, "sd_after.").items():           # BN running-stat side effect
        assert rel_l2(sd[name].double(), v.double()) < 1e-5, name
    with torch.no_grad():
        off_e, _ = align.spatial_transformer(sd, "", g["moving"], g["fixed"], training=False)
    assert rel_l2(off_e, g["offset_eval"]) < TOL


def test_warp_out_of_range():
    g = load_golden("warp")
    img = g["img"].clone().requires_grad_(True)
    grid = g["grid"].clone().requires_grad_(True)
    out = align.warp(img, grid)
    assert rel_l2(out, g["out"]) < 1e-6
    (out * g["w"]).sum().backward()
    assert rel_l2(img.grad, g["g_img"]) < 1e-5
    assert rel_l2(grid.grad, g["g_grid"]) < 1e-4


@pytest.mark.parametrize("tag", ["s", "l"])
def test_losses(tag):
    g = load_golden(f"losses_{tag}")
    fns = dict(ssim=losses.ssimloss, lncc=losses.lncc_loss, mslncc=losses.ms_lncc_loss,
               mi=losses.mi_loss, msmi=losses.ms_mi_loss)
    for name, fn in fns.items():
        if name not in g:
            continue
        X = g["X"].clone().requires_grad_(True)
        Y = g["Y"].clone().requires_grad_(True)
        v = fn(X, Y)
        v.backward()
        assert abs(v.item() - g[name].item()) < 1e-5 * max(1.0, abs(g[name].item())), name
        assert rel_l2(X.grad, g["gX_" + name]) < 1e-4, name
        assert rel_l2(Y.grad, g["gY_" + name]) < 1e-4, name
    if "gauss" in g:
        assert rel_l2(losses.gaussian_smooth(g["X"], 3), g["gauss"]) < 1e-6


def test_rec_step_end_to_end():
    g = load_golden("rec_step")
    sdT = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k)
           for k, v in sub(g, "sdT.").items()}
    sdR = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sub(g, "sdR.").items()}
    inp = step.set_input(g["full"], g["aux"], g["pruned"])
    assert rel_l2(inp["img_k_sampled"], g["k_sampled"]) < 1e-6
    assert rel_l2(inp["img_sampled"], g["img_sampled"]) < 1e-6
    out = step.rec_step(sdT, sdR, inp, g["pruned"], 32, 0.25, num_cascades=2,
                        sens_pools=2, pools=2)
    for k in ("img_offset", "img_warped", "img_rec"):
        assert rel_l2(out[k], g[k]) < TOL, k
    for k in ("loss_all", "loss_smooth", "loss_sim"):
        assert abs(out[k].item() - g[k].item()) < 2e-5 * max(1e-3, abs(g[k].item())), k
    out["loss_all"].backward()
    fl = grad_floor(sub(g, "gT."))
    for name, gg in sub(g, "gT.").items():
        assert rel_l2(sdT[name].grad, gg, fl) < 5e-4, name
    fl = grad_floor(sub(g, "gR."))
    for name, gg in sub(g, "gR.").items():
        assert rel_l2(sdR[name].grad, gg, fl) < 5e-4, name
