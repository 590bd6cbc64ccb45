"""Host logic of the fused NetG / NetD / alignment-U-Net walkers against the reference's golden vectors, on CPU:
the C-ABI ops are replaced by the torch stand-ins of tests/emulation.py (which state each op's documented
semantics), so what is tested is the graph the walkers build - concat order, BatchNorm slices of concatenated
inputs, the space-to-depth weight layout of ConvDown, up-sampling folded into the consumer, buffer updates."""
import torch

import emulation
from conftest import grad_floor, load_golden, rel_l2, sub

TOL = 2e-5


def test_netG_netD_walkers(monkeypatch):
    from spatialalignmentnetwork_b200 import gan
    emulation.install(monkeypatch)
    g = load_golden("gan_s")
    G, D = gan.NetG(1, 1, (4, 8, 12, 8)), gan.NetD(2, ([4] * 2, [8] * 2, [8] * 2))
    G.load_state_dict(sub(g, "sdG."))
    D.load_state_dict(sub(g, "sdD."))
    G.train(); D.train()
    x1, x2 = g["x1"].clone().requires_grad_(True), g["x2"].clone().requires_grad_(True)
    y1, y2 = G(x1), G(x2)
    assert rel_l2(y1, g["y1"]) < TOL and rel_l2(y2, g["y2"]) < TOL
    d1 = D.forward_sources([y1, torch.zeros_like(y1)])
    assert rel_l2(d1, g["d1"]) < TOL
    l_g = gan.loss_gan(d1, real=False, D_loss=False)
    l_1 = (y2 - g["tg"]).abs().mean()
    assert abs(l_g.item() - g["l_g"].item()) < 1e-5
    (l_1 + 0.1 * l_g).backward()
    assert rel_l2(x1.grad, g["g_x1"]) < 5e-4 and rel_l2(x2.grad, g["g_x2"]) < 5e-4
    for pre, net in (("gG.", G), ("gD.", D)):
        ref = sub(g, pre)
        fl = grad_floor(ref)
        for name, p in net.named_parameters():
            assert rel_l2(p.grad, ref[name], fl) < 5e-4, pre + name
    D.zero_grad()
    lf = gan.loss_gan(D.forward_sources([y1.detach(), torch.zeros_like(y1)]), real=False, D_loss=True)
    xr = g["xr"] * 3 - 1
    lr = gan.loss_gan(D(torch.cat([xr, torch.zeros_like(xr)], 1)), real=True, D_loss=True)
    assert abs(lf.item() - g["lf"].item()) < 1e-5 and abs(lr.item() - g["lr"].item()) < 1e-5
    (lf + lr).backward()
    ref = sub(g, "gD2.")
    fl = grad_floor(ref)
    for name, p in D.named_parameters():
        assert rel_l2(p.grad, ref[name], fl) < 5e-4, "gD2." + name
    for pre, net in (("sdG_after.", G), ("sdD_after.", D)):
        sd = net.state_dict()
        for name, ref_v in sub(g, pre).items():
            if name.endswith("num_batches_tracked"):
                assert int(sd[name]) == int(ref_v), name
            else:
                assert rel_l2(sd[name], ref_v) < 1e-5, pre + name
    G.eval(); D.eval()
    with torch.no_grad():
        ye = G(g["x1"])
        de = D.forward_sources([ye, torch.zeros_like(ye)])
    assert rel_l2(ye, g["y_eval"]) < TOL and rel_l2(de, g["d_eval"]) < TOL


def test_alignment_unet_walker(monkeypatch):
    """The same stand-ins under the (GPU-verified) SpatialTransformer / alignment U-Net walker reproduce the
    reference: training forward + backward, running statistics and batch counters afterwards, eval forward."""
    from spatialalignmentnetwork_b200 import cross, model as M, unet as U
    emulation.install(monkeypatch)
    monkeypatch.setattr(U, "USE_TC", True)
    g = load_golden("align_s")
    st = cross.SpatialTransformer(1)
    st.load_state_dict(sub(g, "sd."))
    st.train()
    img = g["img"].clone().requires_grad_(True)
    offset, grid = st(g["moving"], g["fixed"])
    warped = st.warp(img, grid)
    assert rel_l2(offset, g["offset"]) < 1e-4 and rel_l2(grid, g["grid"]) < 1e-5 and rel_l2(warped, g["warped"]) < 1e-4
    ls = M.gradient_loss(offset)
    loss = ((warped - g["tgt"]) ** 2).mean() + 1000.0 * ls
    assert abs(ls.item() - g["loss_smooth"].item()) < 1e-4 * abs(g["loss_smooth"].item())
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * abs(g["loss"].item())
    loss.backward()
    assert rel_l2(img.grad, g["g_img"]) < 1e-3
    from conftest import assert_grads_kink_tolerant
    assert_grads_kink_tolerant({k: p.grad for k, p in st.named_parameters()}, sub(g, "g."), 2e-2, "net_T ")
    sd = st.state_dict()
    for name, ref_v in sub(g, "sd_after.").items():
        if name.endswith("num_batches_tracked"):
            assert int(sd[name]) == int(ref_v), name
        else:
            assert rel_l2(sd[name], ref_v) < 1e-4, name
    st.eval()
    with torch.no_grad():
        off_e, _ = st(g["moving"], g["fixed"])
    assert rel_l2(off_e, g["offset_eval"]) < 1e-4
