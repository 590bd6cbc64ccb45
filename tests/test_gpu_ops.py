"""Per-kernel parity of the san_b200 CUDA ops (through the C ABI) against the CPU oracle /
the same arithmetic in torch fp64 on CPU.  Bar: <= 1e-3 relative (north star); the fp32
kernels are held to much tighter bounds written next to each assert.  Needs a GPU."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _ops():
    from spatialalignmentnetwork_b200 import ops
    return ops


def crandn(*shape):
    return torch.complex(torch.randn(*shape), torch.randn(*shape))


# ------------------------------------------------------------------------------- FFT
@pytest.mark.parametrize("shape", [(2, 3, 16, 20), (1, 2, 20, 23), (1, 1, 320, 320), (2, 1, 64, 368),
                                   (1, 2, 640, 368), (3, 1, 30, 45), (1, 1, 1, 8), (2, 1, 8, 1)])
def test_fft2_ifft2(shape):
    from spatialalignmentnetwork_b200 import signal_utils as su
    torch.manual_seed(0)
    x = crandn(*shape)
    ref_f = torch.fft.fft2(x.to(torch.complex128), norm="ortho")
    ref_i = torch.fft.ifft2(x.to(torch.complex128), norm="ortho")
    xc = x.cuda()
    assert rel_l2(su.fft2(xc), ref_f) < 2e-6
    assert rel_l2(su.ifft2(xc), ref_i) < 2e-6
    assert rel_l2(su.ifft2(su.fft2(xc)), x) < 3e-6          # round trip


def test_fft_golden_and_rss():
    from spatialalignmentnetwork_b200 import signal_utils as su
    for tag in "abc":
        g = load_golden(f"signal_{tag}")
        x = g["x"].cuda()
        assert rel_l2(su.fft2(x), g["fft2"]) < 2e-6
        assert rel_l2(su.rss(x), g["rss"]) < 1e-6
        if tag != "c":
            assert rel_l2(su.ifft2(x), g["ifft2"]) < 2e-6
            assert torch.equal(su.fftshift2(x).cpu(), g["fftshift2"])
            assert torch.equal(su.ifftshift2(x).cpu(), g["ifftshift2"])
            assert rel_l2(su.rss(x.real.contiguous()), g["rss_real"]) < 1e-6


def test_fft_full_size_properties():
    """BASELINE config-2 size: Parseval, linearity and round trip on [64,1,320,320]."""
    from spatialalignmentnetwork_b200 import signal_utils as su
    torch.manual_seed(1)
    x = crandn(64, 1, 320, 320).cuda()
    y = crandn(64, 1, 320, 320).cuda()
    X = su.fft2(x)
    e0, e1 = (x.abs() ** 2).sum().item(), (X.abs() ** 2).sum().item()
    assert abs(e0 - e1) / e0 < 1e-5
    assert rel_l2(su.fft2(x + 2 * y), X + 2 * su.fft2(y)) < 3e-6
    assert rel_l2(su.ifft2(X), x) < 3e-6


def test_fft_autograd():
    from spatialalignmentnetwork_b200 import signal_utils as su
    torch.manual_seed(2)
    x = crandn(2, 2, 20, 16)
    w = crandn(2, 2, 20, 16)
    xr = x.clone().requires_grad_(True)
    (torch.fft.fft2(xr, norm="ortho") * w).real.sum().backward()
    xc = x.cuda().requires_grad_(True)
    (su.fft2(xc) * w.cuda()).real.sum().backward()
    assert rel_l2(xc.grad, xr.grad) < 3e-6


def test_dc_block_golden_and_where_semantics():
    ops = _ops()
    g = load_golden("dc_block")
    k, k0, S, m = g["k"].cuda(), g["k0"].cuda(), g["S"].cuda(), g["mask"].cuda()
    w = g["dc_weight"].cuda()
    red = ops.FftReduce.apply(k, S)
    redc = torch.complex(red[:, :1], red[:, 1:])
    assert rel_l2(redc, g["reduce"]) < 3e-6
    x = redc * (0.5 + 0.25j)
    xp = torch.cat([x.real, x.imag], 1).contiguous()
    out = ops.FftExpandDC.apply(xp, S, k, k0, m, w)
    assert rel_l2(out, g["out"]) < 3e-6
    # torch.where semantics: NaN in k0 at a NOT-sampled column must not leak (bit-exact select)
    k0n = k0.clone()
    col = int((~g["mask"]).nonzero()[0, 0])
    k0n[..., col] = complex(float("nan"), float("nan"))
    out2 = ops.FftExpandDC.apply(xp, S, k, k0n, m, w)
    assert torch.equal(out2, out)


def test_dc_block_backward():
    ops = _ops()
    torch.manual_seed(3)
    N, C, H, W = 2, 3, 16, 20
    k, k0, S = crandn(N, C, H, W), crandn(N, C, H, W), crandn(N, C, H, W)
    xp = torch.randn(N, 2, H, W)
    m = torch.rand(W) > 0.5
    w = torch.tensor([0.7])
    G = crandn(N, C, H, W)

    def ref(k, S, xp, w, k0):
        x = torch.complex(xp[:, :1], xp[:, 1:])
        soft = torch.where(m, k - k0, torch.zeros(1, 1, 1, 1, dtype=k.dtype)) * w
        out = k - soft - torch.fft.fft2(x * S, norm="ortho")
        red = (torch.fft.ifft2(k, norm="ortho") * S.conj()).sum(1, keepdim=True)
        return out, red

    a = [t.clone().double().requires_grad_(True) if not t.is_complex() else t.clone().to(torch.complex128).requires_grad_(True)
         for t in (k, S, xp, w, k0)]
    out, red = ref(*a)
    ((out * G.to(torch.complex128).conj()).real.sum() + (red.real * 0.3 + red.imag * 0.7).sum()).backward()
    b = [t.clone().cuda().requires_grad_(True) for t in (k, S, xp, w, k0)]
    out_c = ops.FftExpandDC.apply(b[2], b[1], b[0], b[4], m.cuda(), b[3])
    red_c = ops.FftReduce.apply(b[0], b[1])
    ((out_c * G.cuda().conj()).real.sum() + (red_c[:, 0] * 0.3 + red_c[:, 1] * 0.7).sum()).backward()
    # k0 = the acquired k-space (VarNetBlock's ref_kspace): d/dk0 = +where(mask, G, 0) * dc_weight
    for name, x, y in zip(("k", "S", "x", "dc_weight", "k0"), b, a):
        assert rel_l2(x.grad, y.grad) < 1e-5, name


@pytest.mark.parametrize("C", [1, 2])
def test_dc_block_320(C):
    """The benchmark size (320 x 320): C = 1 runs the TMA-staged column pass (fft_cols_tma_kernel), C = 2 the register
    kernels with the coil loop; forward, NaN-in-unsampled-column select, and every gradient vs torch complex128
    (reference varnet.py:508-512, 525-530)."""
    ops = _ops()
    torch.manual_seed(11)
    N, H, W = 2, 320, 320
    k, k0, S = crandn(N, C, H, W), crandn(N, C, H, W), crandn(N, C, H, W)
    xp = torch.randn(N, 2, H, W)
    m = torch.rand(W) > 0.6
    m[:8] = False                       # one whole column tile without a sampled column (k0 tile not fetched)
    w = torch.tensor([0.7])
    G = crandn(N, C, H, W)

    def ref(k, S, xp, w, k0):
        x = torch.complex(xp[:, :1], xp[:, 1:])
        soft = torch.where(m, k - k0, torch.zeros(1, 1, 1, 1, dtype=k.dtype)) * w
        out = k - soft - torch.fft.fft2(x * S, norm="ortho")
        red = (torch.fft.ifft2(k, norm="ortho") * S.conj()).sum(1, keepdim=True)
        return out, red

    a = [t.clone().double().requires_grad_(True) if not t.is_complex() else t.clone().to(torch.complex128).requires_grad_(True)
         for t in (k, S, xp, w, k0)]
    out, red = ref(*a)
    ((out * G.to(torch.complex128).conj()).real.sum() + (red.real * 0.3 + red.imag * 0.7).sum()).backward()
    b = [t.clone().cuda().requires_grad_(True) for t in (k, S, xp, w, k0)]
    out_c = ops.FftExpandDC.apply(b[2], b[1], b[0], b[4], m.cuda(), b[3])
    red_c = ops.FftReduce.apply(b[0], b[1])
    assert rel_l2(out_c, out) < 3e-6
    assert rel_l2(torch.complex(red_c[:, :1], red_c[:, 1:]), red) < 3e-6
    ((out_c * G.cuda().conj()).real.sum() + (red_c[:, 0] * 0.3 + red_c[:, 1] * 0.7).sum()).backward()
    for name, x, y in zip(("k", "S", "x", "dc_weight", "k0"), b, a):
        assert rel_l2(x.grad, y.grad) < 1e-5, name
    k0n = k0.clone()
    k0n[..., ~m] = complex(float("nan"), float("nan"))
    with torch.no_grad():
        out2 = ops.FftExpandDC.apply(b[2], b[1], b[0], k0n.cuda(), m.cuda(), b[3])
    assert torch.equal(out2, out_c)


def test_fft_rss_and_sens_normalize():
    ops = _ops()
    torch.manual_seed(4)
    N, C, H, W = 2, 3, 16, 24
    k = crandn(N, C, H, W)
    kr = k.clone().to(torch.complex128).requires_grad_(True)
    r = torch.linalg.vector_norm(torch.fft.ifft2(kr, norm="ortho"), 2, dim=1, keepdim=True)
    wt = torch.rand(N, 1, H, W)
    (r * wt).sum().backward()
    kc = k.cuda().requires_grad_(True)
    rc = ops.FftRss.apply(kc)
    (rc * wt.cuda()).sum().backward()
    assert rel_l2(rc, r) < 3e-6 and rel_l2(kc.grad, kr.grad) < 1e-5
    # sensitivity normalisation
    s = torch.randn(N * C, 2, H, W)
    sr = s.clone().double().requires_grad_(True)
    sc = torch.complex(sr[:, 0], sr[:, 1]).reshape(N, C, H, W)
    Sr = sc / (torch.linalg.vector_norm(sc, 2, dim=1, keepdim=True) + 1e-6)
    Gw = crandn(N, C, H, W)
    (Sr * Gw.to(torch.complex128).conj()).real.sum().backward()
    sg = s.cuda().requires_grad_(True)
    Sg = ops.SensNormalize.apply(sg, N, C)
    (Sg * Gw.cuda().conj()).real.sum().backward()
    assert rel_l2(Sg, Sr) < 2e-6 and rel_l2(sg.grad, sr.grad) < 1e-5


# ------------------------------------------------------------------------------- conv
CONV_CASES = [
    # N, Cin, H, W, Cout, K, bias
    (2, 3, 32, 48, 18, 3, False), (2, 18, 32, 48, 18, 3, False), (1, 36, 20, 20, 72, 3, False),
    (2, 5, 17, 23, 7, 3, True), (2, 32, 40, 40, 64, 1, True), (1, 144, 20, 20, 288, 3, False),
    (2, 18, 16, 16, 2, 1, True), (1, 96, 24, 40, 32, 3, True), (3, 2, 8, 8, 8, 3, True),
    (1, 2, 80, 80, 32, 3, True), (2, 36, 160, 160, 18, 3, False), (1, 64, 10, 12, 128, 1, True),
    (1, 32, 33, 65, 2, 3, True), (2, 72, 40, 40, 36, 3, False),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_bwd(case):
    ops = _ops()
    N, Cin, H, W, Cout, K, has_bias = case
    torch.manual_seed(5)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, K, K) / math.sqrt(Cin * K * K)
    b = torch.randn(Cout) if has_bias else None
    gy = torch.randn(N, Cout, H, W)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    br = b.double().requires_grad_(True) if has_bias else None
    yr = F.conv2d(xr, wr, br, padding=K // 2)
    (yr * gy.double()).sum().backward()
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bc = b.cuda().requires_grad_(True) if has_bias else None
    yc = ops.Conv2d.apply(xc, wc, bc)
    (yc * gy.cuda()).sum().backward()
    assert rel_l2(yc, yr) < 2e-6
    assert rel_l2(xc.grad, xr.grad) < 2e-6
    assert rel_l2(wc.grad, wr.grad) < 1e-5
    if has_bias:
        assert rel_l2(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("case", [(2, 8, 10, 12, 4), (1, 288, 20, 20, 144), (2, 36, 16, 24, 18)])
def test_conv_transpose2x2(case):
    ops = _ops()
    N, Cin, H, W, Cout = case
    torch.manual_seed(6)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cin, Cout, 2, 2) / math.sqrt(Cin)
    gy = torch.randn(N, Cout, 2 * H, 2 * W)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, stride=2)
    (yr * gy.double()).sum().backward()
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    yc = ops.conv_transpose2x2(xc, wc)
    (yc * gy.cuda()).sum().backward()
    assert rel_l2(yc, yr) < 2e-6 and rel_l2(xc.grad, xr.grad) < 2e-6 and rel_l2(wc.grad, wr.grad) < 1e-5


# ------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("shape,slope", [((2, 5, 16, 24), 0.2), ((3, 18, 20, 20), 0.2), ((2, 1, 33, 17), 1.0),
                                         ((1, 4, 320, 320), 0.2)])
def test_instance_norm_lrelu(shape, slope):
    ops = _ops()
    torch.manual_seed(7)
    y = torch.randn(*shape) * 3 + 1.5
    g = torch.randn(*shape)
    yr = y.double().requires_grad_(True)
    outr = F.leaky_relu(F.instance_norm(yr, eps=1e-5), slope)
    (outr * g.double()).sum().backward()
    yc = y.cuda().requires_grad_(True)
    outc = ops.InstanceNormLReLU.apply(yc, slope, 1e-5)
    (outc * g.cuda()).sum().backward()
    assert rel_l2(outc, outr) < 2e-6
    assert rel_l2(yc.grad, yr.grad) < 2e-5


@pytest.mark.parametrize("training", [True, False])
def test_batch_norm_lrelu(training):
    ops = _ops()
    torch.manual_seed(8)
    N, C, H, W = 3, 6, 12, 20
    y = torch.randn(N, C, H, W) * 2 + 0.5
    g = torch.randn(N, C, H, W)
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    rm, rv = torch.randn(C) * 0.1, torch.rand(C) + 0.5
    yr, gr, br = (t.double().requires_grad_(True) for t in (y, gamma, beta))
    rmr, rvr = rm.double().clone(), rv.double().clone()
    outr = F.leaky_relu(F.batch_norm(yr, rmr, rvr, gr, br, training, 0.1, 1e-5), 0.01)
    (outr * g.double()).sum().backward()
    yc, gc, bc = (t.cuda().requires_grad_(True) for t in (y, gamma, beta))
    rmc, rvc = rm.cuda(), rv.cuda()
    outc = ops.BatchNormLReLU.apply(yc, gc, bc, rmc, rvc, training, 0.1, 1e-5, 0.01)
    (outc * g.cuda()).sum().backward()
    assert rel_l2(outc, outr) < 2e-6
    assert rel_l2(yc.grad, yr.grad) < 2e-5
    assert rel_l2(gc.grad, gr.grad) < 1e-5 and rel_l2(bc.grad, br.grad) < 1e-5
    assert rel_l2(rmc, rmr) < 1e-6 and rel_l2(rvc, rvr) < 1e-6


@pytest.mark.parametrize("kind", ["instance", "batch"])
def test_norm_large_mean_over_std(kind):
    """mean/std = 500: the centred form (y - mu) * rstd must keep forward AND backward accurate
    (an a*y + b formulation loses ~mean/std ulps, which showed up as 5e-3 gradient error in net_T)."""
    ops = _ops()
    torch.manual_seed(12)
    N, C, H, W = 2, 4, 16, 24
    y = torch.randn(N, C, H, W) * 0.1 + 50.0
    g = torch.randn(N, C, H, W)
    yr = y.double().requires_grad_(True)
    yc = y.cuda().requires_grad_(True)
    if kind == "instance":
        outr = F.leaky_relu(F.instance_norm(yr, eps=1e-5), 0.2)
        outc = ops.InstanceNormLReLU.apply(yc, 0.2, 1e-5)
    else:
        gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.2
        outr = F.leaky_relu(F.batch_norm(yr, None, None, gamma.double(), beta.double(), True, 0.1, 1e-5), 0.01)
        outc = ops.BatchNormLReLU.apply(yc, gamma.cuda(), beta.cuda(), torch.zeros(C).cuda(), torch.ones(C).cuda(),
                                        True, 0.1, 1e-5, 0.01)
    (outr * g.double()).sum().backward()
    (outc * g.cuda()).sum().backward()
    # the input itself carries 50/0.1 * 2^-24 = 3e-5 relative rounding of (y - mean); allow 4x that
    assert rel_l2(outc, outr) < 1.2e-4
    assert rel_l2(yc.grad, yr.grad) < 1.2e-4


def test_plane_stats_affine_normunet_norm():
    """NormUnet.norm / unnorm (varnet.py:257-273) incl. the gradient through mean and std."""
    from spatialalignmentnetwork_b200.varnet import NormUnet
    torch.manual_seed(9)
    x = torch.randn(3, 2, 16, 24) * 2 + 0.7
    g = torch.randn(3, 2, 16, 24)
    xr = x.double().requires_grad_(True)
    b, c, h, w = xr.shape
    xg = xr.reshape(b, 2, h * w)
    mean, std = xg.mean(2).view(b, 2, 1, 1), xg.std(2).view(b, 2, 1, 1)
    xn = (xr - mean) / (std + 1e-6)
    outr = (xn * xn) * std + mean          # a non-linear "network" between norm and unnorm
    (outr * g.double()).sum().backward()
    nu = NormUnet(4, 2)
    xc = x.cuda().requires_grad_(True)
    xnc, m_, s_ = nu.norm(xc)
    outc = nu.unnorm(xnc * xnc, m_, s_)
    (outc * g.cuda()).sum().backward()
    assert rel_l2(xnc, xn) < 2e-6 and rel_l2(outc, outr) < 2e-6
    assert rel_l2(xc.grad, xr.grad) < 2e-5


def test_pool_up_shuffle_add():
    ops = _ops()
    torch.manual_seed(10)
    x = torch.randn(2, 5, 12, 20)
    g = torch.randn(2, 5, 6, 10)
    xr = x.double().requires_grad_(True)
    pr = F.avg_pool2d(xr, 2)
    (pr * g.double()).sum().backward()
    xc = x.cuda().requires_grad_(True)
    pc = ops.AvgPool2.apply(xc)
    (pc * g.cuda()).sum().backward()
    assert rel_l2(pc, pr) < 1e-6 and rel_l2(xc.grad, xr.grad) < 1e-6
    g2 = torch.randn(2, 5, 24, 40)
    xr = x.double().requires_grad_(True)
    ur = xr.repeat_interleave(2, 2).repeat_interleave(2, 3)
    (ur * g2.double()).sum().backward()
    xc = x.cuda().requires_grad_(True)
    uc = ops.Upsample2.apply(xc)
    (uc * g2.cuda()).sum().backward()
    assert torch.equal(uc.cpu().double(), ur.detach()) and rel_l2(xc.grad, xr.grad) < 1e-6
    a, b_ = torch.randn(7, 3, 5, 5).cuda(), torch.randn(7, 3, 5, 5).cuda()
    assert torch.equal(ops.add(a, b_), a + b_)


# ------------------------------------------------------------------------------- alignment
def test_warp_golden():
    ops = _ops()
    g = load_golden("warp")
    img = g["img"].cuda().requires_grad_(True)
    grid = g["grid"].cuda().requires_grad_(True)
    out = ops.Warp.apply(img, grid)
    assert rel_l2(out, g["out"]) < 1e-6
    (out * g["w"].cuda()).sum().backward()
    assert rel_l2(img.grad, g["g_img"]) < 1e-5
    assert rel_l2(grid.grad, g["g_grid"]) < 1e-4


def test_gradient_loss():
    from spatialalignmentnetwork_b200.model import gradient_loss
    torch.manual_seed(11)
    x = torch.randn(2, 2, 17, 23)                        # network layout [N,2,H,W]
    xr = x.double().requires_grad_(True)
    s = xr.permute(0, 2, 3, 1)
    dx = s[:, :, 1:, :] - s[:, :, :-1, :]
    dy = s[:, 1:, :, :] - s[:, :-1, :, :]
    lr = ((dx * dx).mean() + (dy * dy).mean()) / 2
    (lr * 3.0).backward()
    xc = x.cuda().requires_grad_(True)
    lc = gradient_loss(xc.permute(0, 2, 3, 1))           # non-contiguous view, like cross.py:27-28
    (lc * 3.0).backward()
    assert abs(lc.item() - lr.item()) < 1e-6 * abs(lr.item())
    assert rel_l2(xc.grad, xr.grad) < 1e-5


# ------------------------------------------------------------------------------- losses
@pytest.mark.parametrize("tag", ["s", "l"])
def test_losses_golden(tag):
    from spatialalignmentnetwork_b200 import lnccloss, miloss, ssimloss
    g = load_golden(f"losses_{tag}")
    fns = dict(ssim=ssimloss.ssimloss, lncc=lnccloss.lncc_loss, mslncc=lnccloss.ms_lncc_loss,
               mi=miloss.mi_loss, msmi=miloss.ms_mi_loss)
    for name, fn in fns.items():
        if name not in g:
            continue
        X = g["X"].cuda().requires_grad_(True)
        Y = g["Y"].cuda().requires_grad_(True)
        v = fn(X, Y)
        v.backward()
        assert abs(v.item() - g[name].item()) < 2e-5 * max(1.0, abs(g[name].item())), name
        assert rel_l2(X.grad, g["gX_" + name]) < 2e-4, name
        assert rel_l2(Y.grad, g["gY_" + name]) < 2e-4, name
    if "gauss" in g:
        assert rel_l2(miloss.gaussian_smooth(g["X"].cuda(), 3), g["gauss"]) < 2e-6


def test_losses_vs_oracle_random_sizes():
    from oracle import losses as ol
    from spatialalignmentnetwork_b200 import lnccloss, ssimloss
    torch.manual_seed(12)
    for shp in ((3, 1, 7, 7), (2, 1, 9, 31), (1, 1, 50, 13)):
        X, Y = torch.rand(*shp), torch.rand(*shp)
        for fo, fc in ((ol.ssimloss, ssimloss.ssimloss), (ol.lncc_loss, lnccloss.lncc_loss)):
            a, b = X.double().requires_grad_(True), Y.double().requires_grad_(True)
            vo = fo(a, b)
            vo.backward()
            c, d = X.cuda().requires_grad_(True), Y.cuda().requires_grad_(True)
            vc = fc(c, d)
            vc.backward()
            assert abs(vc.item() - vo.item()) < 5e-5 * max(1.0, abs(vo.item())), (shp, fo.__name__)
            assert rel_l2(c.grad, a.grad) < 5e-4 and rel_l2(d.grad, b.grad) < 5e-4, (shp, fo.__name__)
