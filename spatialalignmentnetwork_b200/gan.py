"""Drop-in for the reference ``gan.py`` (gan.py:7-139): the modality-translation generator ``NetG``, the
patch discriminator ``NetD`` and ``loss_gan``, on the san_b200 kernels.

Every layer of both networks is the pre-activation block of gan.py:10-28,
``spectral_norm(conv)(ReLU(BatchNorm(x)))``.  The module tree (``Conv`` / ``ConvDown`` inside
``CatSequential`` / ``ResSequential`` / ``Sequential``) and the parameter registration are the reference's,
so ``state_dict`` keys (``...norm_layer.weight``, ``...conv.weight_orig``, ``...conv.weight_u`` ...), default
initialisation and RNG consumption are identical and checkpoints interoperate.  The forward pass does not run
the module tree: it walks it and builds fused tcgen05 convolutions (``tc.fused_conv``):

* a conv operand is the channel concatenation of at most two tensors, each read through
  ``ReLU(BatchNorm(.))`` (batch statistics per tensor = the statistics of its slice of the concatenated
  ``BatchNorm2d``) and optionally through nearest x2 up-sampling (``Up()`` commutes with the per-channel
  normalisation and the ReLU; its statistics are those of the low-resolution map);
* residual sums (``ResSequential``) are materialised once (BatchNorm needs the statistics of the sum);
* the kernel-2 stride-2 ``ConvDown`` is a 1x1 convolution over the space-to-depth re-ordering of the
  activated tensor;
* the spectral-norm power iteration, ``W / sigma`` and their backward are ``ops.SpectralNormWeight``.
"""
from functools import partial

import torch

from . import ops, tc
from .unet import CatSequential, NullModule, ResSequential

_SN_EPS = 1e-12   # torch.nn.utils.spectral_norm default


def Down():
    return torch.nn.AvgPool2d(2)


def Up():
    return torch.nn.Upsample(scale_factor=2, mode="nearest")


class Conv(torch.nn.Module):
    """norm_layer -> activation -> spectral_norm(conv) (reference gan.py:10-28)."""

    def __init__(self, in_channels, out_channels,
                 conv=partial(torch.nn.Conv2d, kernel_size=3, padding=1),
                 act=partial(torch.nn.ReLU, inplace=True),
                 norm_layer=torch.nn.BatchNorm2d,
                 weight_norm=torch.nn.utils.spectral_norm,
                 init=torch.nn.init.xavier_normal_):
        super().__init__()
        self.norm_layer = NullModule() if norm_layer is None else norm_layer(in_channels)
        self.act = NullModule() if act is None else act()
        self.conv = conv(in_channels, out_channels)
        init(self.conv.weight)
        # registers conv.weight_orig / weight_u / weight_v exactly like the reference; the hook torch installs
        # is never triggered because the fused forward reads the parameters directly
        self.conv = weight_norm(self.conv)

    # -- pieces used by the fused walkers ---------------------------------------------------------
    def sn_weight(self):
        """Spectrally normalised filter (one power iteration when training)."""
        c = self.conv
        return ops.SpectralNormWeight.apply(c.weight_orig, c.weight_u, c.weight_v, self.training, _SN_EPS)

    def sources(self, node):
        """node: list of (tensor, up) -> tc.Raw terms reading ReLU(norm(tensor))."""
        assert isinstance(self.act, torch.nn.ReLU), "the fused GAN path implements the reference's ReLU blocks"
        bn = self.norm_layer
        if isinstance(bn, NullModule):
            return [tc.Raw(t, None, 0.0, up=up) for t, up in node]
        assert isinstance(bn, torch.nn.BatchNorm2d) and bn.affine and bn.momentum is not None
        out, c0 = [], 0
        for t, up in node:
            C = t.shape[1]
            out.append(tc.Raw(t, "bn", 0.0, bn=_BNSlice(bn, c0, C), up=up))
            c0 += C
        assert c0 == bn.num_features, (c0, bn.num_features)
        return out

    def fused(self, node):
        """Raw fp32 output of the block for a concatenated input ``node``."""
        c = self.conv
        assert c.kernel_size in ((3, 3), (1, 1)) and c.stride == (1, 1)
        return tc.fused_conv(self.sources(node), self.sn_weight(), c.bias)

    def forward(self, x):
        return self.fused([(x.float(), False)])


class ConvDown(Conv):
    """Pre-activation kernel-2 stride-2 convolution (reference gan.py:43-46)."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw, conv=partial(torch.nn.Conv2d, kernel_size=2, stride=2))

    def fused(self, node):
        (t, up), = node
        assert not up
        bn, c = self.norm_layer, self.conv
        if bn.training and bn.track_running_stats:
            bn.num_batches_tracked.add_(1)
        training = bn.training or not bn.track_running_stats
        a = ops.BatchNormLReLU.apply(t, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                     bn.momentum, bn.eps, 0.0)
        s = ops.SpaceToDepth2.apply(a)                                   # [N, 4C, H/2, W/2], channel = c*4 + a*2 + b
        w = self.sn_weight()
        return tc.fused_conv([tc.Raw(s)], w.reshape(w.shape[0], -1, 1, 1), c.bias)


class _BNSlice:
    """Channels [c0, c0 + C) of a ``BatchNorm2d`` as the object ``tc.Raw`` expects: parameter / buffer views
    (the kernels update the running statistics through the views), the batch counter only once."""

    def __init__(self, bn, c0, C):
        self.weight, self.bias = bn.weight[c0:c0 + C], bn.bias[c0:c0 + C]
        self.running_mean = bn.running_mean[c0:c0 + C] if bn.running_mean is not None else None
        self.running_var = bn.running_var[c0:c0 + C] if bn.running_var is not None else None
        self.eps, self.momentum = bn.eps, bn.momentum
        self.training, self.track_running_stats = bn.training, bn.track_running_stats
        self.num_batches_tracked = (bn.num_batches_tracked if c0 == 0 else
                                    torch.zeros_like(bn.num_batches_tracked)) if bn.track_running_stats else None


def _res(res, t):
    """ResSequential (unet.py:15-24) of pre-activation blocks on a materialised tensor: t + subnet(t)."""
    assert res.sample is None
    out = t
    for blk in res.subnet:
        out = blk.fused([(out, False)])
    return ops.add(t, out)


def _cat(cat, t):
    """CatSequential (unet.py:6-13) of NetG -> node [(module(t), up=True), (t, False)]."""
    m = cat.module
    r = _res(m[1], m[0].fused([(t, False)]))
    if len(m) > 3:
        r = _res(m[4], m[3].fused(_cat(m[2], r)))
    assert isinstance(m[len(m) - 1], torch.nn.Upsample)
    return [(r, True), (t, False)]


class NetG(torch.nn.Module):
    """U-Net generator of pre-activation blocks (reference gan.py:72-111)."""

    def __init__(self, in_channels, out_channels, layers):
        super().__init__()
        layers = list(layers)
        num_convs = 2
        current_layer = layers.pop()
        upper_layer = layers.pop()
        unet = CatSequential(
            ConvDown(upper_layer, current_layer),
            ResSequential(*[Conv(current_layer, current_layer) for _ in range(num_convs)]),
            Up())
        for layer in reversed(layers):
            lower_layer, current_layer, upper_layer = current_layer, upper_layer, layer
            unet = CatSequential(
                ConvDown(upper_layer, current_layer),
                ResSequential(*[Conv(current_layer, current_layer) for _ in range(num_convs)]),
                unet,
                Conv(current_layer + lower_layer, current_layer),
                ResSequential(*[Conv(current_layer, current_layer) for _ in range(num_convs - 1)]),
                Up())
        lower_layer, current_layer = current_layer, upper_layer
        self.unet = torch.nn.Sequential(
            Conv(in_channels, current_layer),
            ResSequential(*[Conv(current_layer, current_layer) for _ in range(num_convs - 1)]),
            unet,
            Conv(current_layer + lower_layer, current_layer),
            ResSequential(*[Conv(current_layer, current_layer) for _ in range(num_convs - 1)]),
            Conv(current_layer, out_channels))

    def forward(self, x):
        u = self.unet
        t = _res(u[1], u[0].fused([(x.float(), False)]))
        t = _res(u[4], u[3].fused(_cat(u[2], t)))
        return u[5].fused([(t, False)])


class NetD(torch.nn.Module):
    """Patch discriminator: blocks of un-normalised pre-activation convs separated by 2x2 average pooling, the
    last pooling replaced by the 1-channel head (reference gan.py:113-129)."""

    def __init__(self, in_channels, layers):
        super().__init__()
        out_channels = 1
        layers = list(layers)
        current_layer = in_channels
        conv = partial(Conv, norm_layer=None)
        net = []
        for block in layers:
            for layer in block:
                last_layer, current_layer = current_layer, layer
                net.append(conv(last_layer, current_layer))
            net.append(Down())
        net[-1] = conv(layer, out_channels)
        self.net = torch.nn.Sequential(*net)

    def forward(self, x):
        return self.forward_sources([x])

    def forward_sources(self, images):
        """images: fp32 NCHW tensors whose channel concatenation is the input (no concat copy)."""
        node = [(im.float(), False) for im in images]
        for m in self.net:
            if isinstance(m, torch.nn.AvgPool2d):
                (t, _), = node
                node = [(ops.AvgPool2.apply(t), False)]       # ReLU(pool(y)): the pooling precedes the activation
            else:
                node = [(m.fused(node), False)]
        return node[0][0]


def loss_gan(predict, real=True, D_loss=True):
    """Hinge GAN terms (reference gan.py:131-137)."""
    assert not (real is True and D_loss is False), "are you sure?"
    if D_loss:
        return ops.PairLoss.apply(predict, None, 1, -1.0 if real else 1.0)
    return ops.PairLoss.apply(predict, None, 2, 1.0 if real else -1.0)
