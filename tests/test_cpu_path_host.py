"""Host logic of the VarNet / alignment / CSModel modules against the reference's golden vectors, on CPU: the
C-ABI ops are replaced by the torch stand-ins of tests/emulation.py (each states the documented semantics of the
op it replaces), so what is tested here is the orchestration - the fused U-Net walker (pooling / pixel shuffle /
skip concat order), NormUnet's norm / pad-to-16 / unnorm, the hoisted reference normalisation, the sensitivity
estimator's ACS mask, the cascade loop, set_input, the loss bookkeeping of ``update()``."""
import random

import pytest
import torch

import emulation
from conftest import assert_grads_kink_tolerant, grad_floor, load_golden, rel_l2, sub

TOL = 2e-5


@pytest.mark.parametrize("tag", ["varnet_s", "varnet_p"])
def test_varnet_host_logic(monkeypatch, tag):
    """(1) fp32 against the reference's dump: forward to 2e-5, gradients within the kink-flip criterion (forward
    agreement at 1e-7 still lets a pre-activation at ~0 take the other LeakyReLU branch: 3e-3 on these tiny nets,
    gone with any change of rounding); (2) fp64 against the oracle: the graph the walkers build is EXACTLY the
    reference's (1e-10 on every gradient)."""
    from oracle import varnet as ov
    from spatialalignmentnetwork_b200 import varnet as V
    emulation.install(monkeypatch)
    monkeypatch.setattr(V, "USE_TC", True)
    g = load_golden(tag)
    nc, ch, pools, sch, sp = [int(v) for v in g["cfg"]]
    nlf = int(g["nlf"])
    net = V.VarNet(num_cascades=nc, sens_chans=sch, sens_pools=sp, chans=ch, pools=pools, use_ref=True)
    net.load_state_dict(sub(g, "sd."))
    ks = g["kspace"].clone().requires_grad_(True)
    ref = g["ref"].clone().requires_grad_(True)
    rec = net(ks, ~g["pruned"], ref, nlf)
    assert rel_l2(rec, g["rec"]) < TOL
    assert rel_l2(net.sens_net(g["kspace"], nlf), g["sens"]) < TOL
    ((rec - g["tgt"]) ** 2).mean().backward()
    ours = {"g_kspace": ks.grad, "g_ref": ref.grad, **{k: p.grad for k, p in net.named_parameters()}}
    gold = {"g_kspace": g["g_kspace"], "g_ref": g["g_ref"], **sub(g, "g.")}
    assert_grads_kink_tolerant(ours, gold, 2e-2, tag + ": ")
    # fp64: walkers vs oracle
    net.zero_grad()
    net.double()
    ks = g["kspace"].to(torch.complex128).requires_grad_(True)
    ref = g["ref"].double().requires_grad_(True)
    rec = net(ks, ~g["pruned"], ref, nlf)
    ((rec - g["tgt"].double()) ** 2).mean().backward()
    sd = {k: (v.double() if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point())
          for k, v in sub(g, "sd.").items()}
    ks_o = g["kspace"].to(torch.complex128).requires_grad_(True)
    ref_o = g["ref"].double().requires_grad_(True)
    rec_o = ov.varnet(sd, "", ks_o, ~g["pruned"], ref_o, nlf, nc, sp, pools, use_ref=True)
    ((rec_o - g["tgt"].double()) ** 2).mean().backward()
    assert rel_l2(rec, rec_o) < 1e-12
    assert rel_l2(ks.grad, ks_o.grad) < 1e-10 and rel_l2(ref.grad, ref_o.grad) < 1e-10
    for name, p in net.named_parameters():
        assert rel_l2(p.grad, sd[name].grad, 1e-12) < 1e-10, name


def test_varnet_layerwise_equals_fused_walker(monkeypatch):
    """The two forward implementations of varnet.Unet (layer by layer / fused sources) build the same graph."""
    from spatialalignmentnetwork_b200 import ops, varnet as V
    emulation.install(monkeypatch)
    monkeypatch.setattr(ops, "Conv2d", emulation._Apply(
        lambda x, w, b: torch.nn.functional.conv2d(x, w, b, padding=w.shape[-1] // 2)))
    monkeypatch.setattr(ops, "DepthToSpace2", emulation._Apply(lambda x: torch.nn.functional.pixel_shuffle(x, 2)))
    torch.manual_seed(3)
    u = V.Unet(3, 2, chans=4, num_pool_layers=3)
    x = torch.randn(2, 3, 32, 48)
    monkeypatch.setattr(V, "USE_TC", True)
    a = u(x)
    monkeypatch.setattr(V, "USE_TC", False)
    b = u(x)
    assert rel_l2(a, b) < 1e-5


def test_rec_step_host_logic(monkeypatch):
    """CSModel.set_input / forwardT / forwardR / loss bookkeeping (reg='Rec') against the reference CSModel's dump."""
    from spatialalignmentnetwork_b200 import model as M, unet as U, varnet as V
    emulation.install(monkeypatch)
    monkeypatch.setattr(V, "USE_TC", True)
    monkeypatch.setattr(U, "USE_TC", True)
    g = load_golden("rec_step")
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg="Rec", mask="equispaced", weight_smooth=1000.0,
                   weight_gan=0.1, weight_gan_sim=1.0, weight_sim=1.0, use_amp=False, num_cascades=2, fused_adamw=False,
                   gan_layers_G=[4, 8], gan_layers_D=[[4, 4]])
    random.seed(11)
    net = M.CSModel(cfg)
    net.net_R = V.VarNet(num_cascades=2, sens_chans=2, sens_pools=2, chans=4, pools=2, use_ref=True)
    assert torch.equal(net.net_mask.pruned, g["pruned"])
    net.net_T.load_state_dict(sub(g, "sdT."))
    net.net_R.load_state_dict(sub(g, "sdR."))
    net.train()
    net.set_input(g["full"], g["aux"])
    assert rel_l2(net.img_k_sampled, g["k_sampled"]) < 1e-6 and rel_l2(net.img_sampled, g["img_sampled"]) < 1e-6
    net.loss_all = 0
    net.forwardT()
    net.forwardR()
    for k in ("img_offset", "img_warped", "img_rec"):
        assert rel_l2(getattr(net, k), g[k]) < TOL, k
    for k in ("loss_all", "loss_smooth", "loss_sim"):
        assert abs(getattr(net, k).item() - g[k].item()) < 2e-5 * max(1e-3, abs(g[k].item())), k
    net.loss_all.backward()
    for pre, mod in (("gT.", net.net_T), ("gR.", net.net_R)):
        assert_grads_kink_tolerant({k: p.grad for k, p in mod.named_parameters()}, sub(g, pre), 2e-2, pre)
    vis = net.get_vis("scalars")["scalars"]
    assert {"loss_all", "loss_smooth", "loss_sim"} <= set(vis)


def test_mixed_step_host_logic(monkeypatch):
    """forwardG / forwardD of CSModel (reg='Mixed', model.py:123-190) against the reference CSModel's dump."""
    from spatialalignmentnetwork_b200 import model as M, unet as U, varnet as V
    emulation.install(monkeypatch)
    monkeypatch.setattr(V, "USE_TC", True)
    monkeypatch.setattr(U, "USE_TC", True)
    g = load_golden("mixed_step")
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg="Mixed", mask="equispaced", weight_smooth=1000.0,
                   weight_gan=0.1, weight_gan_sim=1.0, weight_sim=1.0, use_amp=False, num_cascades=2, fused_adamw=False,
                   gan_layers_G=[4, 8, 12, 8], gan_layers_D=[[4, 4], [8, 8], [8, 8]])
    random.seed(12)
    net = M.CSModel(cfg)
    net.net_R = V.VarNet(num_cascades=2, sens_chans=2, sens_pools=2, chans=4, pools=2, use_ref=True)
    for t in "TRGD":
        getattr(net, "net_" + t).load_state_dict(sub(g, f"sd{t}."))
    net.train()
    net.set_input(g["full"], g["aux"])
    net.loss_all = 0
    net.forwardT(); net.forwardG(); net.forwardR(); net.forwardD(D_loss=False)
    # (the tiny golden NetG amplifies the 1e-6 rounding differences of its input ~20x: tests/bf16x3_model.py)
    for k, bar in (("img_warped", TOL), ("img_synth", TOL), ("img_aligned", 2e-4), ("img_rec", TOL)):
        assert rel_l2(getattr(net, k), g[k]) < bar, k
    for k in ("loss_smooth", "loss_sim", "loss_gan_sim", "loss_gan_G"):
        assert abs(getattr(net, k).item() - g[k].item()) < 2e-5 * max(1e-3, abs(g[k].item())), k
    assert abs(net.loss_all.item() - g["loss_G"].item()) < 2e-5 * abs(g["loss_G"].item())
    net.loss_all.backward()
    # same bars as the GPU test: another fp32 rounding realisation of an ill-conditioned tiny fixture
    # (tests/test_oracle_gan.py::test_bf16x3_error_model_sets_the_gpu_bars)
    for t, bar, fl in (("T", 1.2e-1, 5e-2), ("R", 6e-2, 1e-3), ("G", 1e-1, 5e-2)):
        assert_grads_kink_tolerant({k: p.grad for k, p in getattr(net, "net_" + t).named_parameters()},
                                   sub(g, f"g{t}."), bar, f"g{t}.", floor_frac=fl)
    net.loss_all = 0
    net.forwardD(D_loss=True)
    net.net_D.zero_grad()
    for k in ("loss_gan_Dfake", "loss_gan_Dreal"):
        assert abs(getattr(net, k).item() - g[k].item()) < 2e-4 * max(1e-3, abs(g[k].item())), k
    net.loss_all.backward()
    assert_grads_kink_tolerant({k: p.grad for k, p in net.net_D.named_parameters()}, sub(g, "gD."), 6e-2, "gD.")


def test_optional_registration_terms_wiring(monkeypatch):
    """weight_lncc / weight_mi (BASELINE configs 3 / 5): the terms enter loss_all, appear in get_vis and survive a
    full update() - the host side of tests/test_gpu_models.py::test_optional_registration_terms."""
    from oracle import losses as ol
    from spatialalignmentnetwork_b200 import model as M, unet as U, varnet as V
    emulation.install(monkeypatch)
    monkeypatch.setattr(V, "USE_TC", True)
    monkeypatch.setattr(U, "USE_TC", True)
    monkeypatch.setattr(M, "ms_mi_loss", ol.ms_mi_loss)
    monkeypatch.setattr(M, "lncc_loss", ol.lncc_loss)
    torch.manual_seed(5)
    random.seed(5)
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=64, coils=1, reg="Rec", mask="standard", weight_smooth=1000.0,
                   weight_sim=1.0, weight_gan=0.1, weight_gan_sim=1.0, num_cascades=1, weight_lncc=1.0, weight_mi=0.5,
                   gan_layers_G=[8, 16], gan_layers_D=[[8, 8]], fused_adamw=False)
    net = M.CSModel(cfg)
    with torch.no_grad():
        torch.nn.init.normal_(net.net_T.net[-1].weight, 0, 1e-2)
    net.train()
    full, aux = (torch.rand(2, 1, 64, 64) + 0j).to(torch.complex64), (torch.rand(2, 1, 64, 64) + 0j).to(torch.complex64)
    net.set_input(full, aux)
    net.loss_all = 0
    net.forwardT()
    expect = (net.loss_smooth * 1000.0 + ol.lncc_loss(net.img_full_rss, net.img_warped_rss) * 1.0 +
              ol.ms_mi_loss(net.img_full_rss, net.img_warped_rss) * 0.5).item()
    assert abs(net.loss_all.item() - expect) < 1e-5 * abs(expect)
    before = net.net_T.net[-1].weight.detach().clone()
    net.set_input(full, aux)
    net.update()
    assert not torch.equal(before, net.net_T.net[-1].weight.detach())
    assert {"loss_lncc", "loss_mi", "loss_smooth", "loss_sim"} <= set(net.get_vis("scalars")["scalars"])
    # absent weights -> absent terms (the reference's live path)
    cfg2 = M.Config(sparsity=0.25, lr=1e-4, shape=64, coils=1, reg="Rec", mask="standard", weight_smooth=1000.0,
                    weight_sim=1.0, num_cascades=1, gan_layers_G=[8, 16], gan_layers_D=[[8, 8]], fused_adamw=False)
    net2 = M.CSModel(cfg2)
    net2.train()
    net2.set_input(full, aux)
    net2.loss_all = 0
    net2.forwardT()
    assert not hasattr(net2, "loss_lncc") and not hasattr(net2, "loss_mi")


@pytest.mark.parametrize("reg", ["Rec", "GAN-Only"])
def test_test_method_metrics_and_return_value(monkeypatch, reg):
    """CSModel.test() (model.py:265-286): runs forwardT / forwardG / forwardR in eval mode, fills the five metrics
    and returns -PSNR (-MI for GAN-Only); odd batches are rejected like in the reference (chunk of the batch)."""
    import math
    from oracle import gan as og
    from spatialalignmentnetwork_b200 import model as M, unet as U, varnet as V
    emulation.install(monkeypatch)
    monkeypatch.setattr(V, "USE_TC", True)
    monkeypatch.setattr(U, "USE_TC", True)
    torch.manual_seed(6)
    random.seed(6)
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg=reg, mask="equispaced", weight_smooth=1000.0,
                   weight_sim=1.0, weight_gan=0.1, weight_gan_sim=1.0, num_cascades=1, gan_layers_G=[4, 8],
                   gan_layers_D=[[4, 4]], fused_adamw=False)
    net = M.CSModel(cfg)
    net.eval()
    full = (torch.rand(4, 1, 32, 32) + 0j).to(torch.complex64)
    aux = (torch.rand(4, 1, 32, 32) + 0j).to(torch.complex64)
    net.set_input(full, aux)
    r = net.test()
    m = og.metric_sums(net.img_full_rss, net.img_rec)
    assert abs(net.metric_MSE - m["mse"]) < 1e-12 and abs(net.metric_MAE - m["mae"]) < 1e-12
    assert abs(net.metric_PSNR - 10 * math.log10(1 / m["mse"])) < 1e-9
    assert abs(net.metric_MI - og.metric_mi(net.img_full_rss, net.img_warped_rss)) < 1e-12
    assert abs(net.metric_SSIM - (1 - net.loss_sim.item())) < 1e-12
    assert r == (-net.metric_MI if reg == "GAN-Only" else -net.metric_PSNR)
    assert net.img_aligned.shape == net.img_full_rss.shape and net.img_synth.shape == net.img_full_rss.shape
    assert not any(p.grad is not None for p in net.net_R.parameters())          # no_grad
    vis = net.get_vis()
    assert {"metric_PSNR", "metric_SSIM", "metric_MAE", "metric_MSE", "metric_MI"} <= set(vis["scalars"])
    assert {"img_rec", "img_warped", "img_aligned", "img_synth"} <= set(vis["images"])


def test_checkpointed_cascades_and_reg_none(monkeypatch):
    """Host-side knobs: per-cascade recomputation (``VarNet.checkpoint_cascades``) changes nothing; ``reg='None'``
    (model.py:194-205) trains net_R only, with the alignment network evaluated under no_grad."""
    from spatialalignmentnetwork_b200 import model as M, unet as U, varnet as V
    emulation.install(monkeypatch)
    monkeypatch.setattr(V, "USE_TC", True)
    monkeypatch.setattr(U, "USE_TC", True)
    g = load_golden("varnet_s")
    nc, ch, pools, sch, sp = [int(v) for v in g["cfg"]]
    outs = []
    for ck in (False, True):
        net = V.VarNet(num_cascades=nc, sens_chans=sch, sens_pools=sp, chans=ch, pools=pools, use_ref=True)
        net.load_state_dict(sub(g, "sd."))
        net.checkpoint_cascades = ck
        ks = g["kspace"].clone().requires_grad_(True)
        rec = net(ks, ~g["pruned"], g["ref"], int(g["nlf"]))
        ((rec - g["tgt"]) ** 2).mean().backward()
        outs.append((rec.detach(), ks.grad.clone(), net.cascades[0].dc_weight.grad.clone()))
    for a, b in zip(*outs):
        assert rel_l2(b, a) < 1e-6
    torch.manual_seed(8)
    random.seed(8)
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg="None", mask="equispaced", weight_smooth=1000.0,
                   weight_sim=1.0, weight_gan=0.1, weight_gan_sim=1.0, num_cascades=1, gan_layers_G=[4, 8],
                   gan_layers_D=[[4, 4]], fused_adamw=False)
    net = M.CSModel(cfg)
    net.net_R = V.VarNet(num_cascades=1, sens_chans=2, sens_pools=2, chans=4, pools=2, use_ref=True)
    net.optim_R = torch.optim.AdamW(net.net_R.parameters(), lr=1e-4, weight_decay=0)
    net.train()
    snapT = [p.detach().clone() for p in net.net_T.parameters()]
    snapR = [p.detach().clone() for p in net.net_R.parameters()]
    full = (torch.rand(2, 1, 32, 32) + 0j).to(torch.complex64)
    net.set_input(full, full.flip(-1))
    net.update()
    assert all(torch.equal(a, b.detach()) for a, b in zip(snapT, net.net_T.parameters()))
    assert any(not torch.equal(a, b.detach()) for a, b in zip(snapR, net.net_R.parameters()))
    assert all(p.grad is None for p in net.net_T.parameters())
    with pytest.raises(AssertionError):
        net.cfg.reg = "bogus"
        net.set_input(full, full)
        net.update()
