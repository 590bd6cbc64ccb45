"""Drop-in for the live part of the reference ``unet.py``: ``CatSequential``,
``ResSequential``, ``NullModule``, the ``Conv2d`` / ``Up`` / ``Down`` helpers and the
recursive ``UNet`` of the alignment network (reference unet.py:6-31, 119-189).

Container classes keep the reference's nesting, so ``state_dict`` keys (e.g.
``unet.2.module.0.1.weight``) and default initialisation are identical; the forward
passes run the san_b200 kernels: conv (+bias), fused BatchNorm(batch statistics) +
LeakyReLU(0.01), 2x2 average pooling, nearest up-sampling, residual add.
``Encoder`` / ``Decoder`` / ``ResNet`` of the reference are dead code (only referenced
from commented-out lines) and are not provided.
"""
import os

import torch

from . import ops, tc

_SLOPE = 0.01  # torch.nn.LeakyReLU default
USE_TC = os.environ.get("SAN_TC", "1") != "0"   # 0: layer-by-layer fp32 CUDA-core kernels (debug / A-B)


class CatSequential(torch.nn.Module):
    def __init__(self, *modules, dim=1):
        super().__init__()
        self.module = torch.nn.Sequential(*modules)
        self.dim = dim

    def forward(self, x):
        return torch.cat([self.module(x), x], self.dim)  # module output first (unet.py:13)


class ResSequential(torch.nn.Module):
    def __init__(self, *modules, sample=None):
        super().__init__()
        self.subnet = torch.nn.Sequential(*modules)
        self.sample = sample

    def forward(self, x):
        out = self.subnet(x)
        x = self.sample(x) if self.sample is not None else x
        return ops.add(x, out)


class NullModule(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x):
        return x


class Conv2dB200(torch.nn.Conv2d):
    """``torch.nn.Conv2d`` parameters (same init, same keys) + the library conv kernel."""

    def forward(self, x):
        assert self.stride == (1, 1) and self.kernel_size[0] == self.kernel_size[1]
        assert self.padding == (self.kernel_size[0] // 2,) * 2
        return ops.Conv2d.apply(x, self.weight, self.bias)


def _bn_act(bn, x):
    if bn.training and bn.track_running_stats:
        bn.num_batches_tracked.add_(1)
    training = bn.training or not bn.track_running_stats
    return ops.BatchNormLReLU.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                    bn.momentum, bn.eps, _SLOPE)


class _ConvBNAct(torch.nn.Sequential):
    """[Conv2d(+bias), BatchNorm2d, LeakyReLU] with a fused forward."""

    def forward(self, x):
        conv, bn = self[0], self[1]
        return _bn_act(bn, ops.Conv2d.apply(x, conv.weight, conv.bias))


class _UpBlock(torch.nn.Sequential):
    """[Upsample x2 nearest, Conv2d 1x1, BatchNorm2d, LeakyReLU]."""

    def forward(self, x):
        conv, bn = self[1], self[2]
        # a 1x1 conv commutes with nearest up-sampling: convolve at low resolution (4x fewer
        # MACs), then replicate; BatchNorm statistics over the replicated map are unchanged.
        y = ops.Conv2d.apply(x, conv.weight, conv.bias)
        return _bn_act(bn, ops.Upsample2.apply(y))


class _DownBlock(torch.nn.Sequential):
    """[AvgPool2d(2), Conv2d 1x1, BatchNorm2d, LeakyReLU]."""

    def forward(self, x):
        conv, bn = self[1], self[2]
        return _bn_act(bn, ops.Conv2d.apply(ops.AvgPool2.apply(x), conv.weight, conv.bias))


def Conv2d(in_channels, out_channels):
    kernel_size = 3
    return _ConvBNAct(
        torch.nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2),
        torch.nn.BatchNorm2d(out_channels),
        torch.nn.LeakyReLU(inplace=True))


def Up(in_channels, out_channels):
    return _UpBlock(
        torch.nn.Upsample(scale_factor=(2, 2)),
        torch.nn.Conv2d(in_channels, out_channels, kernel_size=1),
        torch.nn.BatchNorm2d(out_channels),
        torch.nn.LeakyReLU(inplace=True))


def Down(in_channels, out_channels):
    return _DownBlock(
        torch.nn.AvgPool2d(2, stride=2),
        torch.nn.Conv2d(in_channels, out_channels, kernel_size=1),
        torch.nn.BatchNorm2d(out_channels),
        torch.nn.LeakyReLU(inplace=True))


class UNet(torch.nn.Module):
    """Recursive U-Net (reference unet.py:144-189); ``layers`` = widths from top to bottom."""

    def __init__(self, in_channels, out_channels, layers):
        super().__init__()
        layers = list(layers)
        num_convs = 2
        current_layer = layers.pop()
        upper_layer = layers.pop()
        unet = CatSequential(
            Down(upper_layer, current_layer),
            ResSequential(*[Conv2d(current_layer, current_layer) for _ in range(num_convs)]),
            Up(current_layer, current_layer))
        for layer in reversed(layers):
            lower_layer, current_layer, upper_layer = current_layer, upper_layer, layer
            unet = CatSequential(
                Down(upper_layer, current_layer),
                ResSequential(*[Conv2d(current_layer, current_layer) for _ in range(num_convs)]),
                unet,
                Conv2d(current_layer + lower_layer, current_layer),
                ResSequential(*[Conv2d(current_layer, current_layer) for _ in range(num_convs - 1)]),
                Up(current_layer, current_layer))
        lower_layer, current_layer = current_layer, upper_layer
        self.unet = torch.nn.Sequential(
            Conv2d(in_channels, current_layer),
            ResSequential(*[Conv2d(current_layer, current_layer) for _ in range(num_convs - 1)]),
            unet,
            Conv2d(current_layer + lower_layer, current_layer),
            ResSequential(*[Conv2d(current_layer, current_layer) for _ in range(num_convs - 1)]),
            Conv2dB200(current_layer, out_channels, 3, padding=1))

    def forward(self, x):
        if USE_TC:
            return self.forward_sources([x])
        return self.unet(x)

    # ---- fused tcgen05 path -------------------------------------------------------------------------
    # Every activated tensor of the reference graph is a SUM of at most two (raw conv output ->
    # BatchNorm -> LeakyReLU) terms (ResSequential, unet.py:15-24) and every conv input a concatenation
    # of such sums (CatSequential, unet.py:6-13), optionally average-pooled (Down) or nearest-up-sampled
    # (Up).  tc.fused_conv applies all of that while staging the operand of the consuming conv.
    @staticmethod
    def _convbn(blk, srcs, modes=None):
        y = tc.fused_conv(srcs, blk[0].weight, blk[0].bias, modes=modes)
        return tc.Raw(y, "bn", _SLOPE, bn=blk[1])

    @classmethod
    def _res(cls, res, S):
        assert res.sample is None
        out = S
        for blk in res.subnet:
            out = [cls._convbn(blk, [out])]
        return list(S) + out                                   # x + subnet(x)

    @classmethod
    def _cat(cls, cat, S):
        """-> the two concatenated sources [module(x), x] of a CatSequential (module output first)."""
        m = cat.module
        yd = tc.fused_conv([S], m[0][1].weight, m[0][1].bias, modes=[tc.MODE_POOL])        # Down: pool -> 1x1
        r = cls._res(m[1], [tc.Raw(yd, "bn", _SLOPE, bn=m[0][2])])
        if len(m) > 3:
            r = cls._res(m[4], [cls._convbn(m[3], cls._cat(m[2], r))])
        up = m[len(m) - 1]
        # Up: a 1x1 conv commutes with nearest up-sampling -> convolve at low resolution, replicate on read
        yu = tc.fused_conv([r], up[1].weight, up[1].bias)
        return [[tc.Raw(yu, "bn", _SLOPE, bn=up[2], up=True)], S]

    def forward_sources(self, images):
        """images: list of fp32 NCHW tensors whose channel concatenation is the network input."""
        u = self.unet
        a = self._res(u[1], [self._convbn(u[0], [tc.Raw(im) for im in images])])
        r = self._res(u[4], [self._convbn(u[3], self._cat(u[2], a))])
        return tc.fused_conv([r], u[5].weight, u[5].bias)
