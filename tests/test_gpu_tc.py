"""Parity of the tcgen05 BF16x3 convolution path (csrc/conv_tc.cu) through the C ABI: operand staging
(fused norm + LeakyReLU + concat + pool / pixel-shuffle / nearest-up), forward, data gradient.
Reference = torch fp64 of the same op (reference call sites varnet.py:98,116,139-146,176-181;
unet.py:119-140).  Bar: BF16x3 keeps ~2^-16 relative operand error -> 2e-5 relative L2 on a conv."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _lib():
    from spatialalignmentnetwork_b200 import _lib
    return _lib


BF16, F16 = 0, 1      # 16-bit pair formats of the staged operands (include/san_b200.h)


def stage(x, fmt=BF16):
    """fp32 NCHW -> staged hi/lo tensor via san_tc_stage_act (single identity source)."""
    L = _lib()
    N, C, H, W = x.shape
    xs = torch.empty(L.lib().san_tc_staged_act_elems(N, H, W, C), dtype=torch.bfloat16, device=x.device)
    Cpad = (C + 7) // 8 * 8          # ceil(C / 8) channel groups
    L.call("tc_stage_act", xs, N, H, W, Cpad, x, None, None, None, 1.0, C, 0,
           None, None, None, None, 1.0, 0, 0, None, None, None, None, 1.0, 0, 0, fmt)
    return xs


def absmax(x):
    L = _lib()
    out = torch.empty(1, device=x.device)
    L.call("absmax", x, x.numel(), out)
    return out


def stage_dyn(x):
    """fp16 pairs with the dynamic power-of-two scale (gradient operands); -> (staged, absmax scalar)."""
    L = _lib()
    N, C, H, W = x.shape
    xs = torch.empty(L.lib().san_tc_staged_act_elems(N, H, W, C), dtype=torch.bfloat16, device=x.device)
    am = absmax(x)
    term = tc_mod()._StageTerm(y=x.data_ptr(), mu=None, a=None, b=None, slope=1.0, C=C, mode=0, accumulate=0)
    arr = (tc_mod()._StageTerm * 1)(term)
    import ctypes
    L.call("tc_stage_terms", xs, N, H, W, (C + 7) // 8 * 8, ctypes.addressof(arr), 1, F16, am)
    return xs, am


def tc_mod():
    from spatialalignmentnetwork_b200 import tc
    return tc


def conv_tc(x, w, bias=None, dgrad=False, fmt=BF16, dyn=False):
    """fmt: pair format of BOTH staged operands (the operands of an MMA share it); dyn: stage x with the dynamic
    scale (as the data gradient stages dY)."""
    L = _lib()
    N, Cin, H, W = x.shape
    Cout, Cin_w, K, _ = w.shape
    fmt_a = fmt_b = fmt
    am = None
    if dyn:
        xs, am = stage_dyn(x)
    else:
        xs = stage(x, fmt_a)
    ws = torch.empty(L.lib().san_tc_staged_weight_elems(H, W, Cin_w if dgrad else Cout, Cout if dgrad else Cin_w, K),
                     dtype=torch.bfloat16, device=x.device)
    L.call("tc_stage_weights", w, ws, H, W, Cout, Cin_w, K, int(dgrad), fmt_b)
    co = Cin_w if dgrad else Cout
    y = torch.empty(N, co, H, W, dtype=torch.float32, device=x.device)
    L.call("tc_conv", xs, ws, bias, y, N, H, W, Cin, co, K, 0, fmt_a + 2 * fmt_b, am)
    return y


TC_CASES = [
    # N, Cin, H, W, Cout, K, bias      (the layer shapes of the cascade / sens / align U-Nets, scaled down)
    (2, 3, 32, 48, 18, 3, False), (1, 18, 64, 64, 18, 3, False), (2, 36, 40, 24, 36, 3, False),
    (1, 72, 20, 20, 144, 3, False), (1, 144, 20, 20, 288, 3, False), (1, 288, 10, 12, 288, 3, False),
    (2, 18, 32, 32, 2, 1, True), (1, 288, 10, 10, 576, 1, False), (2, 2, 33, 47, 32, 3, True),
    (1, 96, 24, 40, 32, 3, True), (3, 64, 17, 23, 64, 1, True), (1, 18, 320, 320, 18, 3, False),
    (1, 96, 320, 320, 32, 3, True),     # dgrad = 32 -> 96 at full width: needs the output-channel split
]


@pytest.mark.parametrize("case", TC_CASES)
def test_tc_conv_fwd_and_dgrad(case):
    N, Cin, H, W, Cout, K, has_bias = case
    torch.manual_seed(21)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, K, K) / math.sqrt(Cin * K * K)
    b = torch.randn(Cout) if has_bias else None
    gy = torch.randn(N, Cout, H, W)
    xr = x.double().requires_grad_(True)
    yr = F.conv2d(xr, w.double(), b.double() if has_bias else None, padding=K // 2)
    (yr * gy.double()).sum().backward()
    # forward as the path runs it: fp16 pairs on both operands = fp32-class (the bar is 10x below the bf16-pair one)
    y = conv_tc(x.cuda(), w.cuda(), b.cuda() if has_bias else None, fmt=F16)
    e16 = rel_l2(y, yr)
    # operand rounding is gone (22 significant bits); what remains is the tensor core's fp32 accumulation, which
    # truncates: the error grows with K = Cin*k*k (measured 1.8e-6 at K = 324, 4.5e-6 at K = 1296) and stays below the
    # bf16-pair error everywhere (checked below)
    e32 = rel_l2(F.conv2d(x, w, b, padding=K // 2), yr)
    assert e16 < (2.5e-6 if Cin * K * K <= 324 else 1e-5), (e16, e32)
    # data gradient as the path runs it: dY as fp16 pairs with the DYNAMIC scale (here gradients of magnitude 1e-7,
    # far below the fp16 range without it), weights as fp16 pairs
    dx = conv_tc(gy.cuda() * 1e-7, w.cuda(), None, dgrad=True, fmt=F16, dyn=True)
    assert rel_l2(dx, xr.grad * 1e-7) < 1e-5
    # bf16 pairs on both operands (SAN_TC_FMT=bf16, the round-1 arithmetic)
    y = conv_tc(x.cuda(), w.cuda(), b.cuda() if has_bias else None)
    eb = rel_l2(y, yr)
    assert eb < 2e-5 and e16 < eb
    dx = conv_tc(gy.cuda(), w.cuda(), None, dgrad=True)
    assert rel_l2(dx, xr.grad) < 2e-5


STATS_CASES = [
    # N, Cin, H, W, Cout, K, bias, group   (group 4 = the pixel-shuffled transposed-conv output, four sub-planes per plane)
    (3, 18, 64, 64, 18, 3, False, 1), (2, 3, 32, 48, 18, 3, False, 1), (2, 36, 40, 24, 36, 3, False, 1),
    (2, 144, 20, 20, 288, 3, False, 1), (2, 288, 10, 12, 576, 1, False, 4), (5, 72, 20, 20, 144, 3, True, 1),
    (2, 18, 320, 320, 18, 3, False, 1), (70, 18, 32, 32, 36, 3, False, 1),
]


@pytest.mark.parametrize("case", STATS_CASES)
def test_tc_conv_statistics_epilogue(case):
    """san_tc_conv_stats + san_in_stats_from_sums = san_plane_stats_in of the conv output (InstanceNorm2d statistics,
    reference varnet.py:141,178), and the conv output itself is unchanged by the statistics epilogue."""
    N, Cin, H, W, Cout, K, has_bias, group = case
    L = _lib()
    assert L.lib().san_tc_conv_stats_supported(H, W, Cin, Cout, K)
    torch.manual_seed(5)
    x = torch.randn(N, Cin, H, W, device="cuda") + 0.3
    w = (torch.randn(Cout, Cin, K, K) / math.sqrt(Cin * K * K)).cuda()
    b = torch.randn(Cout, device="cuda") if has_bias else None
    xs = stage(x, F16)
    ws = torch.empty(L.lib().san_tc_staged_weight_elems(H, W, Cout, Cin, K), dtype=torch.bfloat16, device="cuda")
    L.call("tc_stage_weights", w, ws, H, W, Cout, Cin, K, 0, F16)
    y0 = torch.empty(N, Cout, H, W, device="cuda")
    L.call("tc_conv", xs, ws, b, y0, N, H, W, Cin, Cout, K, 0, 3, None)
    y1 = torch.empty_like(y0)
    sums = torch.full((2 * N * Cout,), 7.0, device="cuda", dtype=torch.float64)        # the call zeroes it
    L.call("tc_conv_stats", xs, ws, b, y1, N, H, W, Cin, Cout, K, 0, 3, None, sums)
    assert torch.equal(y0, y1)
    planes, P = N * Cout // group, H * W
    st = torch.empty(4, planes, device="cuda")
    L.call("in_stats_from_sums", sums, st[0], st[1], st[2], st[3], planes, group, P, 1e-5)
    ref = torch.empty(4, planes, device="cuda")
    L.call("plane_stats_in", y1, ref[0], ref[1], ref[2], ref[3], planes, group * P, 1e-5)
    yd = y1.double().view(planes, -1)
    mean64, var64 = yd.mean(1), yd.var(1, unbiased=False)
    rstd64 = 1.0 / torch.sqrt(var64 + 1e-5)
    # against fp64 and against the separate statistics pass: mean to 1e-6 of the plane's scale, rstd to 2e-6 relative
    scale = torch.sqrt(var64 + mean64 ** 2)
    assert ((st[0].double() - mean64).abs() / scale).max() < 2e-6
    assert ((st[2].double() - rstd64).abs() / rstd64).max() < 4e-6, ((st[2].double() - rstd64).abs() / rstd64).max()
    assert ((st[2] - ref[2]).abs() / ref[2]).max() < 4e-6
    assert ((st[1].double() - var64 * group * P).abs() / (var64 * group * P + 1e-30)).max() < 1e-5
    assert float(st[3].abs().max()) == 0.0


ROWS_CASES = [
    # N, Cin, H, W, Cout, normalised input, dynamic scale
    (3, 18, 64, 64, 18, True, False), (2, 3, 40, 48, 18, False, False), (2, 16, 32, 64, 8, True, False),
    (2, 18, 33, 47, 3, False, True), (70, 8, 32, 32, 8, True, False), (5, 24, 20, 36, 32, True, False),
    (2, 18, 320, 320, 18, True, False), (2, 18, 320, 320, 18, False, True),
]


@pytest.mark.parametrize("case", ROWS_CASES)
def test_tc_conv_rows_fused_staging(case):
    """san_tc_conv_rows (the conv stages its raw fp32 input itself: row ring in shared memory) against the two-pass path
    san_tc_stage_terms + san_tc_conv_stats on the same input: the TMA-stored staged operand is BIT-identical to the staging
    kernel's, the output agrees to accumulation order, the statistics agree; and against torch fp64
    (reference varnet.py:139-146: InstanceNorm + LeakyReLU of the producer, then the 3x3 conv)."""
    N, Cin, H, W, Cout, normed, dyn = case
    L, tc = _lib(), tc_mod()
    assert L.lib().san_tc_conv_rows_supported(H, W, Cin, Cout, 3)
    torch.manual_seed(31)
    x = torch.randn(N, Cin, H, W, device="cuda") * (1e-7 if dyn else 1.5) + (0.0 if dyn else 0.4)
    w = (torch.randn(Cout, Cin, 3, 3) / math.sqrt(Cin * 9)).cuda()
    mu = a = b = None
    slope = 1.0
    if normed:
        mu = torch.randn(N * Cin, device="cuda") * 0.3
        a = torch.rand(N * Cin, device="cuda") + 0.5
        b = torch.randn(N * Cin, device="cuda") * 0.1
        slope = 0.2
    am = absmax(x) if dyn else None
    # reference path: staging pass + conv with the statistics epilogue
    xs_ref = torch.empty(L.lib().san_tc_staged_act_elems(N, H, W, Cin), dtype=torch.bfloat16, device="cuda")
    tc._stage(xs_ref, N, H, W, (Cin + 7) // 8 * 8, [(x, mu, a, b, slope, Cin, 0, False)], F16, am)
    ws_ref = torch.empty(L.lib().san_tc_staged_weight_elems(H, W, Cout, Cin, 3), dtype=torch.bfloat16, device="cuda")
    L.call("tc_stage_weights", w, ws_ref, H, W, Cout, Cin, 3, 0, F16)
    y_ref = torch.empty(N, Cout, H, W, device="cuda")
    sums_ref = torch.empty(2 * N * Cout, dtype=torch.float64, device="cuda")
    L.call("tc_conv_stats", xs_ref, ws_ref, None, y_ref, N, H, W, Cin, Cout, 3, 0, 3, am, sums_ref)
    # row-ring kernel
    ws = torch.empty(L.lib().san_tc_rows_weight_elems(H, W, Cout, Cin), dtype=torch.bfloat16, device="cuda")
    L.call("tc_stage_weights_rows", w, ws, H, W, Cout, Cin, 0, F16)
    xs = torch.full_like(xs_ref, 3.0)
    y = torch.full_like(y_ref, 7.0)
    sums = torch.full_like(sums_ref, 5.0)
    L.call("tc_conv_rows", x, mu, a, b, slope, am, xs, ws, None, y, sums, N, H, W, Cin, Cout)
    torch.cuda.synchronize()
    assert torch.equal(xs.view(torch.int16), xs_ref.view(torch.int16))
    assert rel_l2(y, y_ref) < 2e-6
    assert ((sums - sums_ref).abs() / (sums_ref.abs() + 1e-30 + 1e-6 * sums_ref.abs().max())).max() < 1e-4
    z = x.double()
    if normed:
        z = a.double().view(N, Cin, 1, 1) * (z - mu.double().view(N, Cin, 1, 1)) + b.double().view(N, Cin, 1, 1)
        z = F.leaky_relu(z, slope)
    assert rel_l2(y, F.conv2d(z, w.double(), padding=1)) < 5e-6
    # without the staged copy and without statistics: same output
    y2 = torch.empty_like(y)
    L.call("tc_conv_rows", x, mu, a, b, slope, am, None, ws, None, y2, None, N, H, W, Cin, Cout)
    assert torch.equal(y2, y)


def test_staged_weight_cache_follows_updates():
    """tc._stage_weights caches the staged form of an nn.Parameter on its version counter: in-place torch updates, the
    library's AdamW (raw-pointer kernel + increment_version) and ``.data`` writes followed by invalidate_weight_cache()
    must all be seen; unchanged weights are not staged again."""
    tc = tc_mod()
    from spatialalignmentnetwork_b200 import optim
    L = _lib()
    torch.manual_seed(9)
    w = torch.nn.Parameter((torch.randn(18, 18, 3, 3) / math.sqrt(162)).cuda())
    x = torch.randn(2, 18, 32, 32, device="cuda")

    def check():
        y = tc.fused_conv([tc.Raw(x)], w)
        assert rel_l2(y, F.conv2d(x.double(), w.detach().double(), padding=1)) < 5e-6
        return y

    check()
    n0 = L.launch_count()
    with torch.no_grad():
        tc.fused_conv([tc.Raw(x)], w)
    n_cached = L.launch_count() - n0
    with torch.no_grad():
        w.mul_(0.5)                      # in-place torch op: version bump
    n0 = L.launch_count()
    with torch.no_grad():
        tc.fused_conv([tc.Raw(x)], w)
    assert L.launch_count() - n0 == n_cached + 1      # exactly one more launch: the re-staging of the changed weight
    check()
    opt = optim.AdamW([w], lr=0.05, weight_decay=0.0)
    for _ in range(2):                   # backward uses the cached data-gradient form as well (x needs no gradient here)
        opt.zero_grad()
        xr = x.clone().requires_grad_(True)
        y = tc.fused_conv([tc.Raw(xr)], w)
        (y * y).sum().backward()
        g_ref = torch.autograd.grad((F.conv2d(xr.double(), w.detach().double(), padding=1) ** 2).sum(), xr)[0]
        assert rel_l2(xr.grad, g_ref) < 1e-5
        opt.step()
        check()
    w.data.mul_(2.0)                     # behind autograd's back
    tc.invalidate_weight_cache()
    check()


def test_tc_stage_roundtrip_and_fused_sources():
    """Staging = concat [InstanceNorm+LReLU(0.2) of a pixel-shuffled source, avg-pooled activated source,
    identity source] with zero border / zero pad channels; un-staging returns hi + lo."""
    L = _lib()
    torch.manual_seed(22)
    N, H, W = 2, 12, 20
    ya = torch.randn(N, 4 * 5, H // 2, W // 2) * 2 + 1       # depth-to-space source, 5 channels
    yb = torch.randn(N, 7, 2 * H, 2 * W) - 0.5               # pooled source, 7 channels
    yc = torch.randn(N, 3, H, W)                             # identity source
    # reference
    a_sp = F.pixel_shuffle(ya.double(), 2)
    ra = F.leaky_relu(F.instance_norm(a_sp, eps=1e-5), 0.2)
    rb = F.avg_pool2d(F.leaky_relu(F.instance_norm(yb.double(), eps=1e-5), 0.2), 2)
    ref = torch.cat([ra, rb, yc.double()], 1)
    # coefficients (mu, a = rstd) per plane
    mua, va = a_sp.mean((2, 3)), a_sp.var((2, 3), unbiased=False)
    mub, vb = yb.double().mean((2, 3)), yb.double().var((2, 3), unbiased=False)
    ca = [t.float().reshape(-1).cuda().contiguous() for t in (mua, 1 / torch.sqrt(va + 1e-5))]
    cb = [t.float().reshape(-1).cuda().contiguous() for t in (mub, 1 / torch.sqrt(vb + 1e-5))]
    C = 15
    xs = torch.empty(L.lib().san_tc_staged_act_elems(N, H, W, C), dtype=torch.bfloat16, device="cuda")
    out = torch.empty(N, C, H, W, device="cuda")
    for fmt, bar in ((F16, 5e-7), (BF16, 2e-5)):          # fp16 pairs carry 22 significant bits, bf16 pairs ~17
        L.call("tc_stage_act", xs, N, H, W, 16,
               ya.cuda(), ca[0], ca[1], None, 0.2, 5, 2,
               yb.cuda(), cb[0], cb[1], None, 0.2, 7, 1,
               yc.cuda(), None, None, None, 1.0, 3, 0, fmt)
        L.call("tc_unstage_act", xs, out, N, C, H, W, fmt)
        assert rel_l2(out, ref) < bar, (fmt, rel_l2(out, ref))
    st = xs[8:-256].view(N, 2, 2, H + 2, W + 2, 8).float()      # [lead 8 | planes | trail 256]
    assert xs[:8].float().abs().max() == 0 and xs[-256:].float().abs().max() == 0
    assert st[:, :, :, 0].abs().max() == 0 and st[:, :, :, -1].abs().max() == 0      # zero border rows
    assert st[:, :, :, :, 0].abs().max() == 0 and st[:, :, :, :, -1].abs().max() == 0  # zero border cols
    assert st[:, :, 1, :, :, 7].abs().max() == 0                                       # pad channel 15


def test_tc_conv_full_size_linearity():
    """At the benchmark size (bs 8 here, 320x320, 18->18): conv(a*x1 + x2) = a*conv(x1) + conv(x2) and
    agreement with the fp32 direct-conv kernel on the same inputs."""
    L = _lib()
    from spatialalignmentnetwork_b200 import ops
    torch.manual_seed(23)
    N, C, H, W = 8, 18, 320, 320
    x1, x2 = torch.randn(N, C, H, W, device="cuda"), torch.randn(N, C, H, W, device="cuda")
    w = torch.randn(C, C, 3, 3, device="cuda") / math.sqrt(C * 9)
    y1, y2, y12 = conv_tc(x1, w), conv_tc(x2, w), conv_tc(0.5 * x1 + x2, w)
    assert rel_l2(y12, 0.5 * y1 + y2) < 2e-5
    yf = ops.Conv2d.apply(x1, w, None)
    assert rel_l2(y1, yf) < 2e-5


def _fused_case(kind):
    """One fused block vs torch fp64: forward, gradients wrt every raw input, the weight and the bias."""
    from spatialalignmentnetwork_b200 import tc
    torch.manual_seed(31)
    N, H, W = 2, 16, 24
    if kind == "in_direct":          # conv(lrelu(IN(y)))
        ys = [torch.randn(N, 18, H, W) * 2 + 0.5]
        def ref(y):
            return F.leaky_relu(F.instance_norm(y[0], eps=1e-5), 0.2)
        def srcs(y):
            return [tc.Raw(y[0], "in", 0.2)], None
        Cin = 18
    elif kind == "in_pool":          # conv(avgpool(lrelu(IN(y))))
        ys = [torch.randn(N, 18, 2 * H, 2 * W) + 0.3]
        def ref(y):
            return F.avg_pool2d(F.leaky_relu(F.instance_norm(y[0], eps=1e-5), 0.2), 2)
        def srcs(y):
            return [tc.Raw(y[0], "in", 0.2)], [tc.MODE_POOL]
        Cin = 18
    elif kind == "d2s_skip":         # conv(cat[lrelu(IN(pixel_shuffle(y4))), lrelu(IN(skip))])
        ys = [torch.randn(N, 4 * 18, H // 2, W // 2) - 0.2, torch.randn(N, 18, H, W) * 1.5]
        def ref(y):
            up = F.leaky_relu(F.instance_norm(F.pixel_shuffle(y[0], 2), eps=1e-5), 0.2)
            return torch.cat([up, F.leaky_relu(F.instance_norm(y[1], eps=1e-5), 0.2)], 1)
        def srcs(y):
            return [tc.Raw(y[0], "in", 0.2, d2s=True), tc.Raw(y[1], "in", 0.2)], None
        Cin = 36
    else:                            # identity sources: conv(cat[x, r])
        ys = [torch.randn(N, 2, H, W), torch.randn(N, 1, H, W)]
        def ref(y):
            return torch.cat(y, 1)
        def srcs(y):
            return [tc.Raw(t) for t in y], None
        Cin = 3
    Cout, K = 18, 3
    w = torch.randn(Cout, Cin, K, K) / math.sqrt(Cin * K * K)
    b = torch.randn(Cout)
    gy = torch.randn(N, Cout, H, W)
    yr = [t.double().requires_grad_(True) for t in ys]
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    outr = F.conv2d(ref(yr), wr, br, padding=1)
    (outr * gy.double()).sum().backward()
    yc = [t.cuda().requires_grad_(True) for t in ys]
    wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    s, modes = srcs(yc)
    outc = tc.fused_conv(s, wc, bc, modes=modes)
    (outc * gy.cuda()).sum().backward()
    assert rel_l2(outc, outr) < 2e-5
    for a, r in zip(yc, yr):
        assert rel_l2(a.grad, r.grad) < 5e-5
    assert rel_l2(wc.grad, wr.grad) < 2e-5 and rel_l2(bc.grad, br.grad) < 2e-5


@pytest.mark.parametrize("kind", ["in_direct", "in_pool", "d2s_skip", "identity"])
def test_fused_conv_block(kind):
    _fused_case(kind)


def test_fused_unet_matches_layerwise_fp32():
    """The fused tcgen05 U-Net against the layer-by-layer fp32 CUDA-core path on the same weights
    (forward 2e-5; gradients within the kink noise documented in test_gpu_models.py)."""
    from spatialalignmentnetwork_b200.varnet import Unet
    torch.manual_seed(32)
    net = Unet(3, 2, 18, 4).cuda()
    x = torch.randn(2, 3, 64, 64, device="cuda")
    g = torch.randn(2, 2, 64, 64, device="cuda")
    res = []
    for fused in (True, False):
        xx = x.clone().requires_grad_(True)
        net.zero_grad()
        out = net.forward_sources([xx]) if fused else net._forward_layerwise(xx)
        (out * g).sum().backward()
        res.append((out.detach(), xx.grad.clone(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    assert rel_l2(res[0][0], res[1][0]) < 1e-4        # 23 layers of BF16x3 vs fp32 rounding
    from conftest import assert_grads_kink_tolerant
    assert_grads_kink_tolerant({"dx": res[0][1], **res[0][2]}, {"dx": res[1][1], **res[1][2]}, 2e-2, "fused vs layerwise: ")


WG_CASES = [
    # N, Cin, H, W, Cout, K, bias
    (2, 3, 32, 48, 18, 3, False), (1, 18, 64, 64, 18, 3, False), (2, 36, 40, 24, 36, 3, False),
    (1, 72, 20, 20, 144, 3, False), (1, 144, 20, 20, 288, 3, False), (1, 288, 16, 20, 288, 3, False),
    (2, 18, 32, 32, 2, 1, True), (1, 288, 16, 16, 576, 1, False), (2, 2, 33, 47, 32, 3, True),
    (1, 96, 24, 40, 32, 3, True), (3, 64, 17, 23, 64, 1, True), (2, 18, 320, 320, 18, 3, False),
]


@pytest.mark.parametrize("case", WG_CASES)
def test_tc_wgrad(case):
    """tcgen05 weight gradient from the staged operands vs torch fp64 conv2d backward."""
    L = _lib()
    N, Cin, H, W, Cout, K, has_bias = case
    assert L.lib().san_tc_wgrad_supported(H, W, Cin, Cout, K) == 1
    torch.manual_seed(41)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, K, K) / math.sqrt(Cin * K * K)
    b = torch.randn(Cout) if has_bias else None
    gy = torch.randn(N, Cout, H, W)
    wr = w.double().requires_grad_(True)
    br = b.double().requires_grad_(True) if has_bias else None
    (F.conv2d(x.double(), wr, br, padding=K // 2) * gy.double()).sum().backward()
    for fmt in (F16, BF16):          # the path: fp16 pairs (dY with the dynamic scale); round-1 arithmetic: bf16 pairs
        xs = stage(x.cuda(), fmt)
        gys, am = stage_dyn(gy.cuda()) if fmt == F16 else (stage(gy.cuda()), None)
        dw = torch.empty(Cout, Cin, K, K, device="cuda")
        db = torch.empty(Cout, device="cuda") if has_bias else None
        L.call("tc_wgrad", gys, xs, dw, db, gy.cuda() if has_bias else None, N, H, W, Cin, Cout, K, 3 * fmt, am)
        bar = 1e-5 if fmt == F16 else 2e-5
        assert rel_l2(dw, wr.grad) < bar, (fmt, rel_l2(dw, wr.grad))
        if has_bias:
            assert rel_l2(db, br.grad) < 1e-5


def test_fused_conv_batchnorm_sum_up_pool():
    """net_T operands (unet.py:6-24,119-140): concat[ nearest-up(lrelu(BN(yu))), lrelu(BN(ya)) + lrelu(BN(yb)) ]
    and avg-pool of a sum, BatchNorm in training mode (batch statistics, running-buffer update), vs torch fp64."""
    from spatialalignmentnetwork_b200 import tc
    torch.manual_seed(51)
    N, H, W, C = 3, 16, 24, 8
    ys = [torch.randn(N, C, H // 2, W // 2) + 0.4, torch.randn(N, C, H, W) * 1.3, torch.randn(N, C, H, W) - 0.2]
    bns = [torch.nn.BatchNorm2d(C) for _ in range(3)]
    for bn in bns:
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    w = torch.randn(12, 2 * C, 3, 3) / math.sqrt(2 * C * 9)
    b = torch.randn(12)
    gy = torch.randn(N, 12, H, W)
    import copy
    bnr = [copy.deepcopy(bn).double().train() for bn in bns]
    yr = [t.double().requires_grad_(True) for t in ys]
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    act = lambda bn, t: F.leaky_relu(bn(t), 0.01)
    up = act(bnr[0], yr[0].repeat_interleave(2, 2).repeat_interleave(2, 3))
    outr = F.conv2d(torch.cat([up, act(bnr[1], yr[1]) + act(bnr[2], yr[2])], 1), wr, br, padding=1)
    (outr * gy.double()).sum().backward()
    bnc = [copy.deepcopy(bn).cuda().train() for bn in bns]
    yc = [t.cuda().requires_grad_(True) for t in ys]
    wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    srcs = [[tc.Raw(yc[0], "bn", 0.01, bn=bnc[0], up=True)],
            [tc.Raw(yc[1], "bn", 0.01, bn=bnc[1]), tc.Raw(yc[2], "bn", 0.01, bn=bnc[2])]]
    outc = tc.fused_conv(srcs, wc, bc)
    (outc * gy.cuda()).sum().backward()
    assert rel_l2(outc, outr) < 2e-5
    for a, r in zip(yc, yr):
        assert rel_l2(a.grad, r.grad) < 5e-5
    assert rel_l2(wc.grad, wr.grad) < 2e-5 and rel_l2(bc.grad, br.grad) < 2e-5
    for a, r in zip(bnc, bnr):
        assert rel_l2(a.weight.grad, r.weight.grad) < 2e-5 and rel_l2(a.bias.grad, r.bias.grad) < 2e-5
        assert rel_l2(a.running_mean, r.running_mean) < 1e-6 and rel_l2(a.running_var, r.running_var) < 1e-6
        assert int(a.num_batches_tracked) == 1
    # pooled sum + bare LeakyReLU term
    yp = [torch.randn(N, C, 2 * H, 2 * W), torch.randn(N, C, 2 * H, 2 * W)]
    w2 = torch.randn(5, C, 1, 1)
    g2 = torch.randn(N, 5, H, W)
    ypr = [t.double().requires_grad_(True) for t in yp]
    o2 = F.conv2d(F.avg_pool2d(F.leaky_relu(ypr[0], 0.01) + ypr[1], 2), w2.double())
    (o2 * g2.double()).sum().backward()
    ypc = [t.cuda().requires_grad_(True) for t in yp]
    o2c = tc.fused_conv([[tc.Raw(ypc[0], None, 0.01), tc.Raw(ypc[1])]], w2.cuda(), None, modes=[tc.MODE_POOL])
    (o2c * g2.cuda()).sum().backward()
    assert rel_l2(o2c, o2) < 2e-5
    for a, r in zip(ypc, ypr):
        assert rel_l2(a.grad, r.grad) < 5e-5


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("hw", [(12, 20), (9, 14), (320, 320), (301, 303)])   # Wy % 4 == 0 (vector paths) and not (scalar
def test_in_bwd_fused_map(mode, hw):                          # fall-back); 320 x 320 and 301 x 303: two CTAs of a cluster per plane
    """san_in_bwd_fused_map (InstanceNorm + LeakyReLU backward, reduce + coefficients + apply in one kernel, gradient read
    in place through the adjoint of the consumer's resampling) vs torch fp64 autograd of varnet.py:141-145 followed by
    avg_pool2d / pixel shuffle / nearest x2, and vs the three-kernel path; the absmax side output = max |dy|."""
    L = _lib()
    torch.manual_seed(50 + mode)
    N, Cy, Ctot, c0, slope = 2, 3, 7, 2, 0.2
    Hy, Wy = hw
    if Hy > 256 and mode in (2, 3):      # the up-sampling sources of the full-size case would be 640 x 640 gradients: keep it small
        Hy, Wy = (Hy + 1) // 2 * 2, (Wy + 1) // 2 * 2
    if mode == 1 and (Hy % 2 or Wy % 2):
        Hy, Wy = Hy + Hy % 2, Wy + Wy % 2
    y = (torch.randn(N, 4 * Cy if mode == 2 else Cy, Hy, Wy) * 1.5 + 0.3)
    yr = y.double().requires_grad_(True)
    if mode == 2:
        z = F.leaky_relu(F.instance_norm(yr.reshape(N, Cy, 4 * Hy, Wy), eps=1e-5), slope).reshape(N, 4 * Cy, Hy, Wy)
        op = F.pixel_shuffle(z, 2)
    else:
        z = F.leaky_relu(F.instance_norm(yr, eps=1e-5), slope)
        op = z if mode == 0 else F.avg_pool2d(z, 2) if mode == 1 else F.interpolate(z, scale_factor=2, mode="nearest")
    dx = torch.randn(N, Ctot, op.shape[2], op.shape[3]) * 1e-5          # gradients live far below 1
    (op * dx[:, c0:c0 + Cy].double()).sum().backward()
    yy = y.double().reshape(N, Cy, -1)
    mu = yy.mean(2).reshape(-1).float().cuda()
    a = (1 / torch.sqrt(yy.var(2, unbiased=False) + 1e-5)).reshape(-1).float().cuda()
    yc, dxc = y.cuda(), dx.cuda()
    dy, am = torch.empty_like(yc), torch.empty(1, device="cuda")
    L.call("in_bwd_fused_map", dxc, Ctot, c0, mode, yc, mu, a, slope, dy, N, Cy, Hy, Wy, am)
    assert rel_l2(dy, yr.grad) < 2e-5
    assert abs(am.item() - dy.abs().max().item()) <= 1e-12
    # the three-kernel path on the same inputs
    planes = N * Cy
    wk = torch.empty(5, planes, device="cuda")
    L.call("act_bwd_reduce_map", dxc, Ctot, c0, mode, yc, mu, a, None, a, slope, wk[0], wk[1], N, Cy, Hy, Wy)
    L.call("in_finalize_bwd", wk[0], wk[1], a, wk[2], wk[3], wk[4], planes, y.numel() // planes)
    dy3, am3 = torch.empty_like(yc), torch.empty(1, device="cuda")
    L.call("act_bwd_apply_map", dxc, Ctot, c0, mode, yc, mu, a, None, slope, wk[2], wk[3], wk[4], dy3, N, Cy, Hy, Wy, am3)
    assert rel_l2(dy, dy3) < 2e-6 and abs(am3.item() - dy3.abs().max().item()) <= 1e-12
