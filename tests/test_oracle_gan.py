"""Pin oracle/gan.py (NetG / NetD / spectral norm / the Mixed step / metrics.py) against the golden vectors
minted from the unmodified reference by tests/golden/make_golden_gan.py.  CPU only."""
import torch

from oracle import gan as ogan, step
from conftest import assert_grads_kink_tolerant, grad_floor, load_golden, rel_l2, sub

TOL = 2e-5
G_LEVELS, D_BLOCKS = 3, (2, 2, 2)      # tests/golden/make_golden_gan.py: G_LAYERS = (4, 8, 12, 8), D_LAYERS = 3 x [c, c]


def _sd(g, prefix):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and ("weight_orig" in k or "bias" in k or
                                                                             k.endswith("norm_layer.weight"))
                else v.clone()) for k, v in sub(g, prefix).items()}


def test_netG_netD_forward_backward_and_buffers():
    g = load_golden("gan_s")
    sdG, sdD = _sd(g, "sdG."), _sd(g, "sdD.")
    x1, x2 = g["x1"].clone().requires_grad_(True), g["x2"].clone().requires_grad_(True)
    y1 = ogan.netG(sdG, "unet.", x1, G_LEVELS)
    y2 = ogan.netG(sdG, "unet.", x2, G_LEVELS)
    assert rel_l2(y1, g["y1"]) < TOL and rel_l2(y2, g["y2"]) < TOL
    d1 = ogan.netD(sdD, "net.", torch.cat([y1, torch.zeros_like(y1)], 1), D_BLOCKS)
    assert rel_l2(d1, g["d1"]) < TOL
    l_g = ogan.loss_gan(d1, real=False, D_loss=False)
    l_1 = torch.nn.functional.l1_loss(y2, g["tg"])
    assert abs(l_g.item() - g["l_g"].item()) < 1e-5 and abs(l_1.item() - g["l_1"].item()) < 1e-5
    (l_1 + 0.1 * l_g).backward()
    assert rel_l2(x1.grad, g["g_x1"]) < 5e-4 and rel_l2(x2.grad, g["g_x2"]) < 5e-4
    for pre, sd in (("gG.", sdG), ("gD.", sdD)):
        ref = sub(g, pre)
        fl = grad_floor(ref)
        for name, gg in ref.items():
            assert rel_l2(sd[name].grad, gg, fl) < 5e-4, pre + name
    # discriminator hinge terms (clamp active on both sides)
    for v in sdD.values():
        v.grad = None
    lf = ogan.loss_gan(ogan.netD(sdD, "net.", torch.cat([y1.detach(), torch.zeros_like(y1)], 1), D_BLOCKS), real=False)
    xr = g["xr"] * 3 - 1
    lr = ogan.loss_gan(ogan.netD(sdD, "net.", torch.cat([xr, torch.zeros_like(xr)], 1), D_BLOCKS), real=True)
    assert abs(lf.item() - g["lf"].item()) < 1e-5 and abs(lr.item() - g["lr"].item()) < 1e-5
    (lf + lr).backward()
    ref = sub(g, "gD2.")
    fl = grad_floor(ref)
    for name, gg in ref.items():
        assert rel_l2(sdD[name].grad, gg, fl) < 5e-4, "gD2." + name
    # buffers after two (G) / three (D) training forwards: running statistics, batch counters, power-iteration vectors
    for pre, sd in (("sdG_after.", sdG), ("sdD_after.", sdD)):
        for name, ref_v in sub(g, pre).items():
            if name.endswith("num_batches_tracked"):
                assert int(sd[name]) == int(ref_v), name
            else:
                assert rel_l2(sd[name], ref_v) < 1e-5, pre + name
    with torch.no_grad():
        ye = ogan.netG(sdG, "unet.", g["x1"], G_LEVELS, training=False)
        de = ogan.netD(sdD, "net.", torch.cat([ye, torch.zeros_like(ye)], 1), D_BLOCKS, training=False)
    assert rel_l2(ye, g["y_eval"]) < TOL and rel_l2(de, g["d_eval"]) < TOL


def test_mixed_step_end_to_end():
    g = load_golden("mixed_step")
    req = lambda k, v: v.is_floating_point() and "running" not in k and "weight_u" not in k and "weight_v" not in k
    sds = {t: {k: v.clone().requires_grad_(req(k, v)) for k, v in sub(g, f"sd{t}.").items()} for t in "TRGD"}
    inp = step.set_input(g["full"], g["aux"], g["pruned"])
    out = ogan.mixed_step(sds["T"], sds["R"], sds["G"], sds["D"], inp, g["pruned"], 32, 0.25, num_cascades=2,
                          g_levels=G_LEVELS, d_blocks=D_BLOCKS, sens_pools=2, pools=2)
    for k in ("img_warped", "img_synth", "img_aligned", "img_rec"):
        assert rel_l2(out[k], g[k]) < TOL, k
    for k in ("loss_smooth", "loss_sim", "loss_gan_sim", "loss_gan_G", "loss_G"):
        assert abs(out[k].item() - g[k].item()) < 2e-5 * max(1e-3, abs(g[k].item())), k
    out["loss_G"].backward()
    for t in "TRG":
        assert_grads_kink_tolerant({k: v.grad for k, v in sds[t].items() if v.requires_grad}, sub(g, f"g{t}."), 2e-3, f"g{t}.")
    for v in sds["D"].values():
        v.grad = None
    d = out["d_side"]()
    for k in ("loss_gan_Dfake", "loss_gan_Dreal", "loss_D"):
        assert abs(d[k].item() - g[k].item()) < 2e-5 * max(1e-3, abs(g[k].item())), k
    d["loss_D"].backward()
    assert_grads_kink_tolerant({k: v.grad for k, v in sds["D"].items() if v.requires_grad}, sub(g, "gD."), 2e-3, "gD.")


def test_metrics():
    g = load_golden("metrics")
    m = ogan.metric_sums(g["gt"], g["pred"])
    for k in ("mse", "mae", "nmse"):
        assert abs(m[k] - g[k].item()) < 1e-6 * abs(g[k].item()), k      # reference reduces in fp32
    # (0-dim fixtures are loaded as fp32 tensors: 1e-7; the per-image array keeps fp64: 1e-12)
    assert abs(ogan.metric_mi(g["gt"], g["pred"]) - g["mi"].item()) < 2e-7
    assert abs(ogan.metric_mi(g["gt"], g["pred"], bins=16) - g["mi_16"].item()) < 2e-7
    for i in range(3):
        assert abs(ogan.metric_mi(g["gt"][i:i + 1], g["pred"][i:i + 1]) - g["mi_each"][i].item()) < 1e-12


def _mixed_under_model(g, perturb_seed=None):
    """The oracle's Mixed step with BF16x3 convolutions; ``perturb_seed``: another rounding realisation (net_T's
    filters scaled by 1 + 1e-5 * N(0,1))."""
    import bf16x3_model
    req = lambda k, v: v.is_floating_point() and "running" not in k and "weight_u" not in k and "weight_v" not in k
    sds = {t: {k: v.clone().requires_grad_(req(k, v)) for k, v in sub(g, f"sd{t}.").items()} for t in "TRGD"}
    if perturb_seed is not None:
        torch.manual_seed(perturb_seed)
        for v in sds["T"].values():
            if v.requires_grad and v.dim() == 4:
                v.data.mul_(1 + 1e-5 * torch.randn_like(v))
    inp = step.set_input(g["full"], g["aux"], g["pruned"])
    with bf16x3_model.patched():
        out = ogan.mixed_step(sds["T"], sds["R"], sds["G"], sds["D"], inp, g["pruned"], 32, 0.25, num_cascades=2,
                              g_levels=G_LEVELS, d_blocks=D_BLOCKS, sens_pools=2, pools=2)
        out["loss_G"].backward()
    return out, sds


def test_bf16x3_error_model_sets_the_gpu_bars():
    """What a CORRECT BF16x3 implementation yields on the Mixed-step fixture (tests/bf16x3_model.py): the tiny
    golden NetG amplifies the 5e-5 operand-rounding error of its input (the warped image) ~20x, so
    ``img_aligned`` lands ~1e-3 from the fp32 reference (measured on the B200: 1.15e-3) while every loss stays
    within 2e-5 and the concatenated gradients within 4e-2 (6e-2 over other rounding realisations).  Tiny gradient
    tensors (sums of cancelling terms, < 1 % of the largest gradient norm) move by more than their own size - on
    the B200 ``gG.unet.0.norm_layer.weight`` came out with the opposite sign - which is why the GPU test floors
    per-tensor errors at 5e-2 of the largest norm.  The bars of
    tests/test_gpu_gan.py::test_mixed_step_vs_reference_golden are ~3x these predictions."""
    from conftest import kink_tolerant_failures
    g = load_golden("mixed_step")
    out, sds = _mixed_under_model(g)
    e = {k: rel_l2(out[k], g[k]) for k in ("img_warped", "img_synth", "img_aligned", "img_rec")}
    assert e["img_warped"] < 1e-4 and e["img_synth"] < 1e-4 and e["img_rec"] < 1e-4, e
    assert 3e-4 < e["img_aligned"] < 1.4e-3, e            # the amplified one: above the generic 3e-4 bar, below 4e-3 / 3
    for k in ("loss_smooth", "loss_sim", "loss_gan_sim", "loss_gan_G", "loss_G"):
        assert abs(out[k].item() - g[k].item()) < 1e-4 * max(1e-3, abs(g[k].item())), k
    for seed in (None, 3):
        if seed is not None:
            out, sds = _mixed_under_model(g, seed)
        for t, bar, gpu_bar in (("T", 6e-2, 1.2e-1), ("R", 3e-2, 6e-2), ("G", 5e-2, 1e-1)):
            ours = {k: v.grad for k, v in sds[t].items() if v.requires_grad}
            glob, bad = kink_tolerant_failures(ours, sub(g, f"g{t}."), gpu_bar, floor_frac=5e-2 if t != "R" else 1e-3)
            assert glob < bar and not bad, (seed, t, glob, bad)
    # ... and with the generic 1e-3 floor this realisation fails on one tiny tensor of G, like the B200 run did
    ours = {k: v.grad for k, v in sds["G"].items() if v.requires_grad}
    _, bad = kink_tolerant_failures(ours, sub(g, "gG."), 1e-1, floor_frac=1e-3)
    assert len(bad) == 1 and bad[0][0].startswith("unet.0.norm_layer"), bad
