"""CPU stand-ins for the C-ABI ops the host modules call (test infrastructure).

``tests/test_cpu_gan_host.py`` / ``tests/test_cpu_path_host.py`` monkeypatch them into
``spatialalignmentnetwork_b200.{tc, ops, varnet, cross}`` so that the HOST logic (which tensors are concatenated /
summed, which BatchNorm slice normalises which source, weight re-orderings, padding, hoisting, buffer updates, the
order of the steps of ``CSModel``) can be checked against the reference's golden vectors without a GPU.  Each stand-in states the documented semantics of the op it replaces (include/san_b200.h,
spatialalignmentnetwork_b200/tc.py) with plain torch CPU calls; none of this is reachable from the package."""
import torch
import torch.nn.functional as F

MODE_DIRECT, MODE_POOL, MODE_D2S, MODE_UP = 0, 1, 2, 3


def _activated(r, up):
    """act(norm(raw)) of a ``Raw``, computed ONCE per raw tensor like ``tc.Raw.coef()`` caches its coefficient
    table (a BatchNorm term consumed by two convolutions updates its running statistics once)."""
    key = "_emu_up" if up else "_emu"
    if not hasattr(r, key):
        x = r.y
        if r.d2s:                               # ConvTranspose2d(2, 2) as 1x1 conv + pixel shuffle, then InstanceNorm
            x = F.pixel_shuffle(x, 2)
        if up:                                  # nn.Upsample(nearest x2) BEFORE the normalisation (unet.py:128-133)
            x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        if r.norm == "in":
            x = F.instance_norm(x, eps=1e-5)
        elif r.norm == "bn":
            bn = r.bn
            training = bn.training or not bn.track_running_stats
            if training and bn.track_running_stats:
                bn.num_batches_tracked.add_(1)
            x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)
        if r.slope != 1.0:
            x = F.leaky_relu(x, r.slope)
        setattr(r, key, x)
    return getattr(r, key)


def _term(r, mode):
    """resample(act(norm(raw))) of one ``tc.Raw`` term, in the reference's order of operations."""
    x = _activated(r, mode == MODE_UP)
    if mode == MODE_POOL:                       # avg_pool2d of the ACTIVATED tensor (varnet.py:98)
        x = F.avg_pool2d(x, 2)
    return x


def fused_conv(sources, weight, bias=None, modes=None, stats=False):
    cols = []
    for i, src in enumerate(sources):
        group = src if isinstance(src, (list, tuple)) else [src]
        acc = None
        for r in group:
            mode = modes[i] if modes is not None else (MODE_D2S if r.d2s else MODE_UP if r.up else MODE_DIRECT)
            t = _term(r, mode)
            acc = t if acc is None else acc + t
        cols.append(acc)
    return F.conv2d(torch.cat(cols, 1), weight, bias, padding=weight.shape[-1] // 2)


class _Apply:
    def __init__(self, fn):
        self.apply = fn


class Raw:
    """Stand-in for ``tc.Raw`` (same fields; any floating dtype, so the walkers can also be run in fp64 for an
    exact graph comparison with the oracle)."""

    def __init__(self, y, norm=None, slope=1.0, d2s=False, bn=None, up=False):
        assert y.dim() == 4 and (norm == "bn") == (bn is not None)
        self.y, self.norm, self.slope, self.d2s, self.bn, self.up = y, norm, float(slope), bool(d2s), bn, bool(up)


def _bn_lrelu(y, gamma, beta, rm, rv, training, momentum, eps, slope):
    return F.leaky_relu(F.batch_norm(y, rm, rv, gamma, beta, training, momentum, eps), slope)


def _sn_weight(w, u, v, power_iteration, eps):
    wm = w.reshape(w.shape[0], -1)
    if power_iteration:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))
    sigma = torch.dot(u.clone(), torch.mv(wm, v.clone()))
    return w / sigma


def _pair_loss(x, y, mode, sign):
    if mode == 0:
        return (x - y).abs().mean()
    if mode == 1:
        return torch.clamp(sign * x, min=-1).mean()
    return (sign * x).mean()


# ---- FFT / data consistency / sensitivity maps (include/san_b200.h, "FFT / data consistency") ----------------
def _planar(c):            # complex [N,1,H,W] -> planar [N,2,H,W]
    return torch.cat([c.real, c.imag], dim=1)


def _cplx(p):              # planar [N,2,H,W] -> complex [N,1,H,W]
    return torch.complex(p[:, :1], p[:, 1:])


def _fft2(x, inverse):
    return (torch.fft.ifft2 if inverse else torch.fft.fft2)(x, norm="ortho")


def _ifft_masked_planar(k, colmask):
    N, C, H, W = k.shape
    img = torch.fft.ifft2(k * colmask.to(k.real.dtype), norm="ortho").reshape(N * C, 1, H, W)
    return _planar(img)


def _fft_reduce(k, sens):
    return _planar((torch.fft.ifft2(k, norm="ortho") * sens.conj()).sum(dim=1, keepdim=True))


def _fft_expand_dc(x, sens, k, k0, mask, dc_weight):
    zero = torch.zeros(1, 1, 1, 1, dtype=k.dtype)
    return k - torch.where(mask, k - k0, zero) * dc_weight - torch.fft.fft2(_cplx(x) * sens, norm="ortho")


def _expand(xp, sens):
    return torch.fft.fft2(_cplx(xp) * sens, norm="ortho")


def _fft_rss(k):
    return torch.linalg.vector_norm(torch.fft.ifft2(k, norm="ortho"), 2, dim=1, keepdim=True)


def _rss(x):
    return torch.linalg.vector_norm(x, 2, dim=1, keepdim=True)


def _sens_normalize(s, N, C):
    _, _, H, W = s.shape
    c = _cplx(s).reshape(N, C, H, W)
    return c / (_rss(c) + 1e-6)


# ---- layer-wise norm / resampling ops ---------------------------------------------------------------------------
def _plane_stats(x):
    N, C, H, W = x.shape
    mean = x.mean(dim=(2, 3))
    # d m2 / dx = 2 (x - mean) exactly (ops.PlaneStats.backward): centring on the detached mean drops the term
    # -2 sum(x - mean) / P, zero in exact arithmetic but rounding noise in fp32
    return mean, ((x - mean.detach()[:, :, None, None]) ** 2).sum(dim=(2, 3))


def _plane_affine(x, mu, a, b):
    N, C = x.shape[:2]
    v = x if mu is None else x - mu.reshape(N, C, 1, 1)
    v = v * a.reshape(N, C, 1, 1)
    return v if b is None else v + b.reshape(N, C, 1, 1)


def _in_lrelu(y, slope, eps):
    return F.leaky_relu(F.instance_norm(y, eps=eps), slope)


# ---- alignment / losses ---------------------------------------------------------------------------------------------
def _grid_from_offset(x):
    N, _, H, W = x.shape
    theta = torch.eye(2, 3, dtype=x.dtype)[None].expand(N, 2, 3)
    return F.affine_grid(theta, (N, 1, H, W), align_corners=False) + x.permute(0, 2, 3, 1)


def _warp(img, grid):
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def _gradient_loss(s):
    dx = s[:, :, 1:, :] - s[:, :, :-1, :]
    dy = s[:, 1:, :, :] - s[:, :-1, :, :]
    return ((dx * dx).mean() + (dy * dy).mean()) / 2.0


def _error_sums(a, b):
    d = a.double() - b.double()
    return (d * d).sum().item(), d.abs().sum().item(), (a.double() ** 2).sum().item()


def install(monkeypatch):
    """Route the op calls of the host modules to the stand-ins above."""
    from oracle import gan as og, losses as ol
    from spatialalignmentnetwork_b200 import cross, ops, tc, varnet
    monkeypatch.setattr(tc, "fused_conv", fused_conv)
    monkeypatch.setattr(tc, "Raw", Raw)
    monkeypatch.setattr(ops, "add", lambda a, b: a + b)
    monkeypatch.setattr(ops, "BatchNormLReLU", _Apply(_bn_lrelu))
    monkeypatch.setattr(ops, "SpaceToDepth2", _Apply(lambda x: F.pixel_unshuffle(x, 2)))
    monkeypatch.setattr(ops, "AvgPool2", _Apply(lambda x: F.avg_pool2d(x, 2)))
    monkeypatch.setattr(ops, "SpectralNormWeight", _Apply(_sn_weight))
    monkeypatch.setattr(ops, "PairLoss", _Apply(_pair_loss))
    monkeypatch.setattr(ops, "l1_loss", lambda x, y: (x - y).abs().mean())
    monkeypatch.setattr(ops, "Fft2", _Apply(_fft2))
    monkeypatch.setattr(ops, "IfftMaskedPlanar", _Apply(_ifft_masked_planar))
    monkeypatch.setattr(ops, "FftReduce", _Apply(_fft_reduce))
    monkeypatch.setattr(ops, "FftExpandDC", _Apply(_fft_expand_dc))
    monkeypatch.setattr(ops, "FftRss", _Apply(_fft_rss))
    monkeypatch.setattr(ops, "Rss", _Apply(_rss))
    monkeypatch.setattr(ops, "SensNormalize", _Apply(_sens_normalize))
    monkeypatch.setattr(ops, "PlaneStats", _Apply(_plane_stats))
    monkeypatch.setattr(ops, "PlaneAffine", _Apply(_plane_affine))
    monkeypatch.setattr(ops, "InstanceNormLReLU", _Apply(_in_lrelu))
    monkeypatch.setattr(ops, "GridFromOffset", _Apply(_grid_from_offset))
    monkeypatch.setattr(ops, "Warp", _Apply(_warp))
    monkeypatch.setattr(ops, "GradientLoss", _Apply(_gradient_loss))
    monkeypatch.setattr(ops, "SsimLoss", _Apply(ol.ssimloss))
    monkeypatch.setattr(ops, "LnccLoss", _Apply(ol.lncc_loss))
    monkeypatch.setattr(ops, "error_sums", _error_sums)
    monkeypatch.setattr(ops, "mi_metric", lambda gt, pred, bins=64, minVal=0.0, maxVal=1.0: og.metric_mi(gt, pred, bins, minVal, maxVal))
    monkeypatch.setattr(varnet, "_Expand", _Apply(_expand))
    monkeypatch.setattr(cross, "_LReLUFn", _Apply(lambda x, slope: F.leaky_relu(x, slope)))
