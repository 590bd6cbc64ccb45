"""GAN branch, device metrics and the multi-tensor AdamW through the C ABI:
op level against the same arithmetic in torch fp64 / the CPU oracle, module level against golden vectors
minted from the unmodified reference (tests/golden/make_golden_gan.py).  Needs a GPU."""
import random

import pytest
import torch
import torch.nn.functional as F

from conftest import assert_grads_kink_tolerant, grad_floor, load_golden, rel_l2, sub

pytestmark = pytest.mark.gpu

TOL_NET = 3e-4    # outputs of whole networks: ~25 BF16x3 convs chained through BatchNorm (cf. TOL_T of test_gpu_models.py)
GTOL_TINY = 6e-2  # whole-network gradients of the tiny golden nets (kink flips, tests/test_gpu_models.py)


# ------------------------------------------------------------------------------------------- ops
@pytest.mark.parametrize("shape", [(8, 4, 3, 3), (12, 20, 3, 3), (512, 64, 2, 2), (1, 8, 3, 3), (256, 512, 3, 3)])
@pytest.mark.parametrize("training", [True, False])
def test_spectral_norm_weight(shape, training):
    """gan.py:24 torch.nn.utils.spectral_norm: value, u / v update, gradient (u, v constants)."""
    from spatialalignmentnetwork_b200 import ops
    torch.manual_seed(1)
    w = torch.randn(*shape)
    u = F.normalize(torch.randn(shape[0]), dim=0)
    v = F.normalize(torch.randn(w.numel() // shape[0]), dim=0)
    gy = torch.randn(*shape)
    wd = w.double().requires_grad_(True)
    wm = wd.reshape(shape[0], -1)
    ud, vd = u.double(), v.double()
    if training:
        with torch.no_grad():
            vd = F.normalize(wm.t() @ ud, dim=0, eps=1e-12)
            ud = F.normalize(wm @ vd, dim=0, eps=1e-12)
    ref = wd / torch.dot(ud, wm @ vd)
    (ref * gy.double()).sum().backward()
    wc, uc, vc = w.cuda().requires_grad_(True), u.cuda(), v.cuda()
    out = ops.SpectralNormWeight.apply(wc, uc, vc, training, 1e-12)
    (out * gy.cuda()).sum().backward()
    assert rel_l2(out, ref) < 2e-6
    assert rel_l2(uc, ud) < 2e-6 and rel_l2(vc, vd) < 2e-6
    assert rel_l2(wc.grad, wd.grad) < 2e-5
    if not training:
        assert torch.equal(uc.cpu(), u) and torch.equal(vc.cpu(), v)


@pytest.mark.parametrize("mode,sign", [(0, 1.0), (1, 1.0), (1, -1.0), (2, -1.0)])
def test_pair_losses(mode, sign):
    """F.l1_loss (model.py:138) and the hinge / linear terms of loss_gan (gan.py:131-137)."""
    from spatialalignmentnetwork_b200 import ops
    torch.manual_seed(2)
    x = (torch.randn(3, 1, 37, 41) * 1.5).requires_grad_(True)
    y = torch.randn(3, 1, 37, 41).requires_grad_(True)
    with torch.no_grad():
        y[0, 0, 0, :4] = x[0, 0, 0, :4]                      # exact ties: sign(0) = 0
    if mode == 0:
        ref = F.l1_loss(x, y)
    elif mode == 1:
        ref = torch.clamp(sign * x, min=-1).mean()
    else:
        ref = (sign * x).mean()
    (ref * 1.7).backward()
    xc, yc = x.detach().cuda().requires_grad_(True), y.detach().cuda().requires_grad_(True)
    out = ops.PairLoss.apply(xc, yc if mode == 0 else None, mode, sign)
    (out * 1.7).backward()
    assert abs(out.item() - ref.item()) < 1e-6 * max(1.0, abs(ref.item()))
    assert rel_l2(xc.grad, x.grad) < 1e-6
    if mode == 0:
        assert rel_l2(yc.grad, y.grad) < 1e-6


def test_loss_gan_semantics():
    from spatialalignmentnetwork_b200 import gan
    from oracle import gan as ogan
    torch.manual_seed(3)
    p = torch.randn(2, 1, 20, 20) * 2
    for real, d_loss in ((True, True), (False, True), (False, False)):
        assert abs(gan.loss_gan(p.cuda(), real=real, D_loss=d_loss).item() - ogan.loss_gan(p, real, d_loss).item()) < 1e-6
    with pytest.raises(AssertionError):
        gan.loss_gan(p.cuda(), real=True, D_loss=False)


def test_space_to_depth_is_stride2_conv():
    """The k2 s2 convolution of gan.py:43-46 == 1x1 convolution of the space-to-depth tensor with the
    reshaped filter."""
    from spatialalignmentnetwork_b200 import ops
    torch.manual_seed(4)
    x = torch.randn(2, 5, 12, 16, requires_grad=True)
    w = torch.randn(7, 5, 2, 2)
    ref = F.conv2d(x, w, stride=2)
    xc = x.detach().cuda().requires_grad_(True)
    s = ops.SpaceToDepth2.apply(xc)
    assert torch.equal(s.cpu(), F.pixel_unshuffle(x.detach(), 2))
    out = F.conv2d(s.cpu(), w.reshape(7, 20, 1, 1))
    assert rel_l2(out, ref) < 1e-6
    g = torch.randn_like(s)
    (s * g).sum().backward()
    assert torch.equal(xc.grad.cpu(), F.pixel_shuffle(g.cpu(), 2))


def test_metrics_against_reference_values():
    """metrics.py mse / mae / nmse / mi on the device against values computed by the reference's functions."""
    from spatialalignmentnetwork_b200 import ops
    g = load_golden("metrics")
    gt, pred = g["gt"].cuda(), g["pred"].cuda()
    se, ae, sa = ops.error_sums(gt, pred)
    n = gt.numel()
    assert abs(se / n - g["mse"].item()) < 1e-6 * g["mse"].item()
    assert abs(ae / n - g["mae"].item()) < 1e-6 * g["mae"].item()
    assert abs(se / sa - g["nmse"].item()) < 1e-6 * g["nmse"].item()
    assert abs(ops.mi_metric(gt, pred) - g["mi"].item()) < 2e-7
    assert abs(ops.mi_metric(gt, pred, bins=16) - g["mi_16"].item()) < 2e-7
    for i in range(3):
        assert abs(ops.mi_metric(gt[i:i + 1], pred[i:i + 1]) - g["mi_each"][i].item()) < 1e-12


def test_mi_metric_full_size_against_oracle():
    from spatialalignmentnetwork_b200 import ops
    from oracle import gan as ogan
    torch.manual_seed(5)
    a = torch.rand(4, 1, 320, 320)
    b = (a * 0.6 + 0.4 * torch.rand(4, 1, 320, 320)) * 1.05 - 0.02           # some samples outside [0, 1]
    assert abs(ops.mi_metric(a.cuda(), b.cuda()) - ogan.metric_mi(a, b)) < 1e-10


def test_adamw_matches_torch():
    """san_adamw_step == torch.optim.AdamW (model.py:72-81 settings, plus weight decay) over several steps,
    including > 48 tensors (several launches), tensors larger than one chunk and a parameter without gradient."""
    from spatialalignmentnetwork_b200.optim import AdamW
    torch.manual_seed(6)
    shapes = [(3,), (17, 5), (70000,), (4, 4, 3, 3)] * 14 + [(300, 301)]
    for wd in (0.0, 0.05):
        p_ref = [torch.randn(*s).cuda().requires_grad_(True) for s in shapes]
        p_our = [p.detach().clone().requires_grad_(True) for p in p_ref]
        o_ref = torch.optim.AdamW(p_ref, lr=1e-2, weight_decay=wd)
        o_our = AdamW(p_our, lr=1e-2, weight_decay=wd)
        for it in range(4):
            for i, (a, b) in enumerate(zip(p_ref, p_our)):
                if i == 2 and it < 2:
                    a.grad = b.grad = None            # joins later: its own step counter
                    continue
                gr = torch.randn_like(a) * (10.0 ** (i % 5 - 2))
                a.grad, b.grad = gr, gr.clone()
            o_ref.step(); o_our.step()
        for a, b in zip(p_ref, p_our):
            assert rel_l2(b, a) < 2e-6
        sr, so = o_ref.state_dict()["state"], o_our.state_dict()["state"]
        assert set(sr) == set(so) and int(so[0]["step"]) == 4 and int(so[2]["step"]) == 2
        assert rel_l2(so[1]["exp_avg_sq"], sr[1]["exp_avg_sq"]) < 2e-6 and rel_l2(so[1]["exp_avg"], sr[1]["exp_avg"]) < 2e-6


# --------------------------------------------------------------------------------------- modules
def test_netG_netD_vs_reference_golden():
    """Same sequence as tests/golden/make_golden_gan.py: two training forwards of G, one of D, backward; the
    discriminator's hinge terms on detached inputs, backward; buffers; an eval forward."""
    from spatialalignmentnetwork_b200 import gan, ops
    g = load_golden("gan_s")
    G, D = gan.NetG(1, 1, (4, 8, 12, 8)), gan.NetD(2, ([4] * 2, [8] * 2, [8] * 2))
    G.load_state_dict(sub(g, "sdG."))
    D.load_state_dict(sub(g, "sdD."))
    G.cuda().train(); D.cuda().train()
    x1, x2 = g["x1"].cuda().requires_grad_(True), g["x2"].cuda().requires_grad_(True)
    y1, y2 = G(x1), G(x2)
    assert rel_l2(y1, g["y1"]) < TOL_NET and rel_l2(y2, g["y2"]) < TOL_NET
    d1 = D(torch.cat([y1, torch.zeros_like(y1)], 1))
    assert rel_l2(d1, g["d1"]) < TOL_NET
    l_g = gan.loss_gan(d1, real=False, D_loss=False)
    l_1 = ops.l1_loss(y2, g["tg"].cuda())
    assert abs(l_g.item() - g["l_g"].item()) < 1e-4 and abs(l_1.item() - g["l_1"].item()) < 1e-4
    (l_1 + 0.1 * l_g).backward()
    assert_grads_kink_tolerant({"x1": x1.grad, "x2": x2.grad}, {"x1": g["g_x1"], "x2": g["g_x2"]}, GTOL_TINY, "inputs ")
    assert_grads_kink_tolerant({k: p.grad for k, p in G.named_parameters()}, sub(g, "gG."), GTOL_TINY, "gG.")
    assert_grads_kink_tolerant({k: p.grad for k, p in D.named_parameters()}, sub(g, "gD."), GTOL_TINY, "gD.")
    D.zero_grad()
    lf = gan.loss_gan(D.forward_sources([y1.detach(), torch.zeros_like(y1)]), real=False, D_loss=True)
    xr = (g["xr"] * 3 - 1).cuda()
    lr = gan.loss_gan(D.forward_sources([xr, torch.zeros_like(xr)]), real=True, D_loss=True)
    assert abs(lf.item() - g["lf"].item()) < 1e-4 and abs(lr.item() - g["lr"].item()) < 1e-4
    (lf + lr).backward()
    assert_grads_kink_tolerant({k: p.grad for k, p in D.named_parameters()}, sub(g, "gD2."), GTOL_TINY, "gD2.")
    # running statistics / batch counters after two training forwards, u / v after two (G) / three (D) power iterations
    for pre, net in (("sdG_after.", G), ("sdD_after.", D)):
        sd = net.state_dict()
        for name, ref_v in sub(g, pre).items():
            if name.endswith("num_batches_tracked"):
                assert int(sd[name]) == int(ref_v), name
            else:
                assert rel_l2(sd[name], ref_v) < 2e-4, pre + name
    G.eval(); D.eval()
    with torch.no_grad():
        ye = G(g["x1"].cuda())
        de = D.forward_sources([ye, torch.zeros_like(ye)])
    assert rel_l2(ye, g["y_eval"]) < TOL_NET and rel_l2(de, g["d_eval"]) < TOL_NET


def _mixed_model(g, reg="Mixed"):
    from spatialalignmentnetwork_b200 import model as M
    from spatialalignmentnetwork_b200.varnet import VarNet
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg=reg, mask="equispaced",
                   weight_smooth=1000.0, weight_gan=0.1, weight_gan_sim=1.0, weight_sim=1.0, use_amp=False,
                   num_cascades=2, gan_layers_G=[4, 8, 12, 8], gan_layers_D=[[4, 4], [8, 8], [8, 8]])
    random.seed(12)
    net = M.CSModel(cfg)
    net.net_R = VarNet(num_cascades=2, sens_chans=2, sens_pools=2, chans=4, pools=2, use_ref=True)
    net.optim_R = type(net.optim_T)(net.net_R.parameters(), lr=cfg.lr, weight_decay=0)
    assert torch.equal(net.net_mask.pruned, g["pruned"])
    for t in "TRGD":
        getattr(net, "net_" + t).load_state_dict(sub(g, f"sd{t}."))
    return net.to("cuda")


def test_mixed_step_vs_reference_golden():
    """CSModel reg='Mixed' (model.py:123-190, 217-239): generator-side pass + backward, then the discriminator
    pass, against the reference CSModel's dump."""
    g = load_golden("mixed_step")
    net = _mixed_model(g)
    net.train()
    net.set_input(g["full"].cuda(), g["aux"].cuda())
    net.loss_all = 0
    net.forwardT(); net.forwardG(); net.forwardR(); net.forwardD(D_loss=False)
    # Bars = 3x what a correct BF16x3 implementation yields on this fixture according to the CPU error model
    # (tests/bf16x3_model.py, tests/test_oracle_gan.py::test_bf16x3_error_model_sets_the_gpu_bars): the tiny
    # golden NetG amplifies the operand-rounding error of the warped image it translates ~20x, so img_aligned
    # lands ~1.1e-3 from the fp32 reference; concatenated gradients 1.2e-2 (R) .. 3.7e-2 (T).
    for k, bar in (("img_warped", 3e-4), ("img_synth", 3e-4), ("img_aligned", 4e-3), ("img_rec", 3e-4)):
        assert rel_l2(getattr(net, k), g[k]) < bar, k
    for k in ("loss_smooth", "loss_sim", "loss_gan_sim", "loss_gan_G"):
        assert abs(getattr(net, k).item() - g[k].item()) < 3e-4 * max(1e-3, abs(g[k].item())), k
    assert abs(net.loss_all.item() - g["loss_G"].item()) < 3e-4 * abs(g["loss_G"].item())
    net.loss_all.backward()
    # Small tensors (e.g. the 1-element BatchNorm affine of NetG's first block, 0.3 % of the largest gradient norm)
    # are sums of cancelling terms: rounding realisations of the CPU model move them by up to 8e-3 of the largest
    # norm, sign included (3 of 5 realisations fail a 1e-3 floor, none a 3e-2 floor:
    # tests/test_oracle_gan.py::test_bf16x3_error_model_sets_the_gpu_bars) -> floor 5e-2 for T and G here.
    for t, bar, fl in (("T", 1.2e-1, 5e-2), ("R", GTOL_TINY, 1e-3), ("G", 1e-1, 5e-2)):
        assert_grads_kink_tolerant({k: p.grad for k, p in getattr(net, "net_" + t).named_parameters()},
                                   sub(g, f"g{t}."), bar, f"g{t}.", floor_frac=fl)
    net.loss_all = 0
    net.forwardD(D_loss=True)
    net.optim_D.zero_grad()
    for k in ("loss_gan_Dfake", "loss_gan_Dreal"):
        assert abs(getattr(net, k).item() - g[k].item()) < 3e-4 * max(1e-3, abs(g[k].item())), k
    assert abs(net.loss_all.item() - g["loss_D"].item()) < 3e-4 * abs(g["loss_D"].item())
    net.loss_all.backward()
    assert_grads_kink_tolerant({k: p.grad for k, p in net.net_D.named_parameters()}, sub(g, "gD."), GTOL_TINY, "gD.")


@pytest.mark.parametrize("reg", ["Mixed", "GAN-Only"])
def test_update_gan_modes_and_test_metrics(reg):
    """update() in the GAN modes steps T, G (, R) and then D; test() fills every metric of model.py:265-286."""
    g = load_golden("mixed_step")
    net = _mixed_model(g, reg)
    net.train()
    snap = {t: [p.detach().clone() for p in getattr(net, "net_" + t).parameters()] for t in "TRGD"}
    net.set_input(g["full"].cuda(), g["aux"].cuda())
    net.update()
    changed = {t: any(not torch.equal(a, b.detach()) for a, b in zip(snap[t], getattr(net, "net_" + t).parameters()))
               for t in "TRGD"}
    assert changed == {"T": True, "G": True, "D": True, "R": reg == "Mixed"}
    sc = net.get_vis("scalars")["scalars"]
    assert {"loss_smooth", "loss_gan_sim", "loss_gan_G", "loss_gan_Dfake", "loss_gan_Dreal"} <= set(sc)
    assert all(v == v for v in sc.values())                       # finite
    net.eval()
    net.set_input(g["full"].cuda(), g["aux"].cuda())
    r = net.test()
    assert r == (-net.metric_MI if reg == "GAN-Only" else -net.metric_PSNR)
    from oracle import gan as ogan
    m = ogan.metric_sums(net.img_full_rss.cpu(), net.img_rec.cpu())
    assert abs(net.metric_MSE - m["mse"]) < 1e-9 and abs(net.metric_MAE - m["mae"]) < 1e-8
    assert abs(net.metric_PSNR - m["psnr"]) < 1e-6
    assert abs(net.metric_MI - ogan.metric_mi(net.img_full_rss.cpu(), net.img_warped_rss.cpu())) < 1e-10
    assert 0.0 <= net.metric_SSIM <= 1.0 and net.img_aligned.shape == net.img_full_rss.shape
