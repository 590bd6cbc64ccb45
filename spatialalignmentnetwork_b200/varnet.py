"""Drop-in for the reference ``varnet.py`` (E2E-VarNet with a reference-modality channel).

Same classes, constructor arguments, ``forward`` signatures and ``state_dict`` keys as
the reference (varnet.py:24-530); the standard ``torch.nn`` layer objects are kept only
as parameter containers (identical default initialisation and checkpoint keys) while
every ``forward`` runs the san_b200 CUDA kernels through ``ops``: fused
FFT + coil-reduce / FFT + soft-DC kernels instead of cuFFT + ATen, and the library's
conv / norm / resampling kernels instead of cuDNN + ATen.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from . import ops, tc
from .signal_utils import rss

# SAN_TC=0 selects the fp32 CUDA-core conv kernels layer by layer (debug / A-B only); the default is
# the fused tcgen05 path of tc.py.
USE_TC = os.environ.get("SAN_TC", "1") != "0"

_IN_EPS = 1e-5


class ConvBlock(nn.Module):
    """2 x [conv3x3 no-bias -> InstanceNorm -> LeakyReLU(0.2)] (reference varnet.py:122-156)."""

    def __init__(self, in_chans: int, out_chans: int):
        super().__init__()
        self.in_chans = in_chans
        self.out_chans = out_chans
        self.layers = nn.Sequential(
            nn.Conv2d(in_chans, out_chans, kernel_size=3, padding=1, bias=False),
            nn.InstanceNorm2d(out_chans),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
            nn.InstanceNorm2d(out_chans),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
        )

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        x = ops.Conv2d.apply(image, self.layers[0].weight, None)
        x = ops.InstanceNormLReLU.apply(x, 0.2, _IN_EPS)
        x = ops.Conv2d.apply(x, self.layers[3].weight, None)
        return ops.InstanceNormLReLU.apply(x, 0.2, _IN_EPS)


class TransposeConvBlock(nn.Module):
    """ConvTranspose2d(2, stride 2, no bias) -> InstanceNorm -> LeakyReLU(0.2)
    (reference varnet.py:159-192)."""

    def __init__(self, in_chans: int, out_chans: int):
        super().__init__()
        self.in_chans = in_chans
        self.out_chans = out_chans
        self.layers = nn.Sequential(
            nn.ConvTranspose2d(in_chans, out_chans, kernel_size=2, stride=2, bias=False),
            nn.InstanceNorm2d(out_chans),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
        )

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        x = ops.conv_transpose2x2(image, self.layers[0].weight)
        return ops.InstanceNormLReLU.apply(x, 0.2, _IN_EPS)


class _Conv1x1(nn.Conv2d):
    """nn.Conv2d(kernel 1) parameter container whose forward is the library kernel."""

    def forward(self, x):
        return ops.Conv2d.apply(x, self.weight, self.bias)


class Unet(nn.Module):
    """Real-valued U-Net (reference varnet.py:24-119)."""

    def __init__(self, in_chans: int, out_chans: int, chans: int = 32, num_pool_layers: int = 4):
        super().__init__()
        self.in_chans = in_chans
        self.out_chans = out_chans
        self.chans = chans
        self.num_pool_layers = num_pool_layers
        self.down_sample_layers = nn.ModuleList([ConvBlock(in_chans, chans)])
        ch = chans
        for _ in range(num_pool_layers - 1):
            self.down_sample_layers.append(ConvBlock(ch, ch * 2))
            ch *= 2
        self.conv = ConvBlock(ch, ch * 2)
        self.up_conv = nn.ModuleList()
        self.up_transpose_conv = nn.ModuleList()
        for _ in range(num_pool_layers - 1):
            self.up_transpose_conv.append(TransposeConvBlock(ch * 2, ch))
            self.up_conv.append(ConvBlock(ch * 2, ch))
            ch //= 2
        self.up_transpose_conv.append(TransposeConvBlock(ch * 2, ch))
        self.up_conv.append(nn.Sequential(ConvBlock(ch * 2, ch), _Conv1x1(ch, self.out_chans, kernel_size=1, stride=1)))

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        assert not torch.is_complex(image)
        if USE_TC:
            return self.forward_sources([image])
        return self._forward_layerwise(image)

    def forward_sources(self, images) -> torch.Tensor:
        """Fused tcgen05 path.  ``images``: list of fp32 NCHW tensors whose channel concatenation is the
        network input (so NormUnet never materialises ``cat([x, ref])``).  Every conv reads its operand
        through ``tc.fused_conv``: InstanceNorm + LeakyReLU(0.2) of the producing layer, the 2x2 average
        pooling (varnet.py:98), the ConvTranspose2d pixel shuffle (:176-179) and the skip concat (:116,
        up-sampled first) are all applied while the operand tiles are staged."""
        R = tc.Raw
        h, w = images[0].shape[-2:]
        assert h % (1 << self.num_pool_layers) == 0 and w % (1 << self.num_pool_layers) == 0, \
            "fused path needs H, W divisible by 2**num_pool_layers (NormUnet pads to 16)"
        srcs, modes = [R(im) for im in images], None
        stack = []
        for layer in self.down_sample_layers:
            y1 = tc.fused_conv(srcs, layer.layers[0].weight, modes=modes, stats=True)
            y2 = R(tc.fused_conv([R(y1, "in", 0.2)], layer.layers[3].weight, stats=True), "in", 0.2)
            stack.append(y2)
            srcs, modes = [y2], [tc.MODE_POOL]
        y1 = tc.fused_conv(srcs, self.conv.layers[0].weight, modes=modes, stats=True)
        cur = R(tc.fused_conv([R(y1, "in", 0.2)], self.conv.layers[3].weight, stats=True), "in", 0.2)
        out = None
        for transpose_conv, conv in zip(self.up_transpose_conv, self.up_conv):
            skip = stack.pop()
            wt = transpose_conv.layers[0].weight                       # [Cin, Cout, 2, 2]
            w1 = wt.permute(1, 2, 3, 0).reshape(wt.shape[1] * 4, wt.shape[0], 1, 1)
            y4 = R(tc.fused_conv([cur], w1, stats=True), "in", 0.2, d2s=True)      # 1x1 conv to 4*Cout; shuffle on read
            block = conv if isinstance(conv, ConvBlock) else conv[0]
            y1 = tc.fused_conv([y4, skip], block.layers[0].weight, stats=True)     # cat([up, skip]) (varnet.py:116)
            cur = R(tc.fused_conv([R(y1, "in", 0.2)], block.layers[3].weight, stats=True), "in", 0.2)
            if not isinstance(conv, ConvBlock):
                out = tc.fused_conv([cur], conv[1].weight, conv[1].bias)   # final 1x1 + bias (varnet.py:78)
        return out

    def _forward_layerwise(self, image: torch.Tensor) -> torch.Tensor:
        stack = []
        output = image
        for layer in self.down_sample_layers:
            output = layer(output)
            stack.append(output)
            output = ops.AvgPool2.apply(output)
        output = self.conv(output)
        for transpose_conv, conv in zip(self.up_transpose_conv, self.up_conv):
            skip = stack.pop()
            output = transpose_conv(output)
            padding = [0, 0, 0, 0]
            if output.shape[-1] != skip.shape[-1]:
                padding[1] = 1
            if output.shape[-2] != skip.shape[-2]:
                padding[3] = 1
            if sum(padding) != 0:  # odd sizes only; never reached after NormUnet's pad-to-16
                output = F.pad(output, padding, "reflect")
            output = torch.cat([output, skip], dim=1)  # up-sampled first (varnet.py:116)
            output = conv(output)
        return output


class NormUnet(nn.Module):
    """Normalised U-Net on complex input (reference varnet.py:200-332)."""

    def __init__(self, chans: int, num_pools: int, in_chans: int = 1, out_chans: int = 1, use_ref: bool = False):
        super().__init__()
        self.use_ref = use_ref
        if self.use_ref:
            self.unet = Unet(in_chans=in_chans * 3, out_chans=out_chans * 2, chans=chans, num_pool_layers=num_pools)
            self.ref_norm = torch.nn.InstanceNorm2d(in_chans)
        else:
            self.unet = Unet(in_chans=in_chans * 2, out_chans=out_chans * 2, chans=chans, num_pool_layers=num_pools)
        self.in_chans = in_chans
        self.out_chans = out_chans

    # complex <-> planar helpers of the reference API (varnet.py:246-255)
    def complex_to_chan_dim(self, x: torch.Tensor) -> torch.Tensor:
        assert torch.is_complex(x)
        return torch.cat([x.real, x.imag], dim=1)

    def chan_complex_to_last_dim(self, x: torch.Tensor) -> torch.Tensor:
        assert not torch.is_complex(x)
        _, c, _, _ = x.shape
        assert c % 2 == 0
        c = c // 2
        return torch.complex(x[:, :c], x[:, c:])

    def norm(self, x: torch.Tensor):
        """Per-sample, per-{re,im} group mean / unbiased std (reference varnet.py:257-268)."""
        assert not torch.is_complex(x)
        b, c, h, w = x.shape
        assert c == 2, "san_b200 NormUnet supports one complex channel (c == 2), as VarNet uses it"
        mean, m2 = ops.PlaneStats.apply(x)
        std = torch.sqrt(m2 / (h * w - 1))
        a = 1.0 / (std + 1e-6)
        xn = ops.PlaneAffine.apply(x, mean, a, None)          # (x - mean) / (std + 1e-6)
        return xn, mean.view(b, 2, 1, 1), std.view(b, 2, 1, 1)

    def unnorm(self, x, mean, std):
        b = x.shape[0]
        return ops.PlaneAffine.apply(x, None, std.reshape(b, 2), mean.reshape(b, 2))   # x * std + mean

    def pad(self, x):
        _, _, h, w = x.shape
        w_mult = ((w - 1) | 15) + 1
        h_mult = ((h - 1) | 15) + 1
        w_pad = [math.floor((w_mult - w) / 2), math.ceil((w_mult - w) / 2)]
        h_pad = [math.floor((h_mult - h) / 2), math.ceil((h_mult - h) / 2)]
        if w_mult != w or h_mult != h:
            x = F.pad(x, w_pad + h_pad)
        return x, (h_pad, w_pad, h_mult, w_mult)

    def unpad(self, x, h_pad, w_pad, h_mult, w_mult):
        if h_pad[0] == 0 and h_pad[1] == 0 and w_pad[0] == 0 and w_pad[1] == 0:
            return x
        return x[..., h_pad[0]:h_mult - h_pad[1], w_pad[0]:w_mult - w_pad[1]].contiguous()

    def normalize_ref(self, ref):
        """InstanceNorm2d of the reference image (varnet.py:235,317); parameter-free, so VarNet
        hoists it out of the cascade loop."""
        assert not torch.is_complex(ref)
        return ops.InstanceNormLReLU.apply(ref, 1.0, _IN_EPS)

    def forward_planar(self, x, ref_normed=None):
        """x: planar [N,2,H,W] float -> planar [N,2,H,W] (the kernels' native layout)."""
        x, mean, std = self.norm(x)
        x, pad_sizes = self.pad(x)
        if self.use_ref:
            assert ref_normed is not None
            r, _ = self.pad(ref_normed)
            x = self.unet.forward_sources([x, r]) if USE_TC else self.unet(torch.cat([x, r], dim=1))
        else:
            assert ref_normed is None
            x = self.unet(x)
        x = self.unpad(x, *pad_sizes)
        return self.unnorm(x, mean, std)

    def forward(self, x: torch.Tensor, ref: torch.Tensor = None) -> torch.Tensor:
        assert len(x.shape) == 4
        assert torch.is_complex(x)
        assert x.shape[1] == self.in_chans
        x = self.complex_to_chan_dim(x).contiguous()
        if self.use_ref:
            assert not torch.is_complex(ref)
            ref = self.normalize_ref(ref)
        else:
            assert ref is None
        x = self.forward_planar(x, ref)
        x = self.chan_complex_to_last_dim(x)
        assert x.shape[1] == self.out_chans
        return x


class SensitivityModel(nn.Module):
    """Coil-sensitivity estimator (reference varnet.py:335-420)."""

    def __init__(self, chans: int, num_pools: int, in_chans: int = 1, out_chans: int = 1, mask_center: bool = True):
        super().__init__()
        self.mask_center = mask_center
        self.norm_unet = NormUnet(chans, num_pools, in_chans=in_chans, out_chans=out_chans)

    def forward(self, masked_kspace: torch.Tensor, num_low_frequencies: int) -> torch.Tensor:
        N, C, H, W = masked_kspace.shape
        # ACS low-pass column mask (reference varnet.py:395-398), built on the device (no host copy: the step stays
        # capturable into a CUDA graph)
        acs = torch.ones(W, device=masked_kspace.device)
        acs[num_low_frequencies:] = 0
        acs = torch.roll(acs, -num_low_frequencies // 2)
        # ifft2(ACS * k) straight into the planar layout; the reference's chunked U-Net
        # (varnet.py:409-414) is a memory workaround with identical per-sample arithmetic
        images = ops.IfftMaskedPlanar.apply(masked_kspace, acs)      # [N*C, 2, H, W]
        s = self.norm_unet.forward_planar(images)
        return ops.SensNormalize.apply(s, N, C)


class VarNetBlock(nn.Module):
    """One cascade: soft data consistency + U-Net regulariser (reference varnet.py:488-530)."""

    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model
        self.dc_weight = nn.Parameter(torch.ones(1))

    def sens_expand(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        xp = torch.cat([x.real, x.imag], dim=1).contiguous()
        return _Expand.apply(xp, sens_maps)

    def sens_reduce(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        xp = ops.FftReduce.apply(x, sens_maps)
        return torch.complex(xp[:, :1], xp[:, 1:])

    def forward(self, current_kspace, ref_kspace, mask, sens_maps, ref_image, _ref_normed=None):
        mask = mask.reshape(-1)
        xp = ops.FftReduce.apply(current_kspace, sens_maps)
        if self.model.use_ref:
            if _ref_normed is None:
                _ref_normed = self.model.normalize_ref(ref_image)
            yp = self.model.forward_planar(xp, _ref_normed)
        else:
            yp = self.model.forward_planar(xp)
        return ops.FftExpandDC.apply(yp, sens_maps, current_kspace, ref_kspace, mask, self.dc_weight)


class _Expand(torch.autograd.Function):
    """fft2(x * S) alone (the reference's ``sens_expand`` helper, varnet.py:508-509)."""

    @staticmethod
    def forward(ctx, xp, sens):
        xp, sens = ops._c(xp), ops._c(sens)
        N, C, H, W = sens.shape
        out = torch.empty_like(sens)
        ops.call("fft_expand_dc", xp, sens, None, None, None, None, out, out, N, C, H, W, 0)
        ctx.save_for_backward(xp, sens)
        return out

    @staticmethod
    def backward(ctx, G):
        xp, sens = ctx.saved_tensors
        N, C, H, W = sens.shape
        G = ops._c(G)
        dx = torch.empty_like(xp)
        u = torch.empty_like(sens)
        tmp = torch.empty_like(sens)
        ops.call("fft_reduce", G, sens, dx, u, tmp, N, C, H, W, 1, 1.0)
        dS = torch.empty_like(sens)
        ops.call("cmul_conj_planar", u, xp, dS, N, C, H * W, 1.0)
        return dx, dS


class VarNet(nn.Module):
    """Full variational network (reference varnet.py:422-486)."""

    def __init__(self, num_cascades: int = 12, sens_chans: int = 8, sens_pools: int = 4, chans: int = 18,
                 pools: int = 4, mask_center: bool = True, use_ref: bool = False):
        super().__init__()
        self.use_ref = use_ref
        self.sens_net = SensitivityModel(chans=sens_chans, num_pools=sens_pools, mask_center=mask_center)
        self.cascades = nn.ModuleList(
            [VarNetBlock(NormUnet(chans, pools, use_ref=use_ref)) for _ in range(num_cascades)])
        self.checkpoint_cascades = False  # recompute each cascade in backward (memory knob, not in the reference)

    def forward(self, masked_kspace, mask, ref, num_low_frequencies):
        ns = N_STREAMS
        if ns > 1 and masked_kspace.is_cuda and masked_kspace.shape[0] >= MIN_SLICES_PER_STREAM * ns:
            return self._forward_streams(masked_kspace, mask, ref, num_low_frequencies, ns)
        return self._forward_one(masked_kspace, mask, ref, num_low_frequencies)

    def _forward_streams(self, masked_kspace, mask, ref, num_low_frequencies, ns):
        """The batch as ``ns`` independent sub-batches, each on its own CUDA stream (every operation of the network is
        per-sample - InstanceNorm, the NormUnet normalisation, the FFTs -, so the result is the un-split one; the weight
        gradients of the sub-batches are summed by autograd).  Why: a step is a strict chain of kernels that alternate
        between tensor / shared-memory-bound convolutions (HBM 30-40 % busy) and HBM-bound element-wise passes (staging,
        statistics, normalisation backward).  Two chains give the GPU a convolution of one sub-batch to run next to the
        element-wise kernels of the other (two tcgen05 kernels never co-reside: the hardware serialises those).  The backward
        follows by itself: autograd runs every backward node on the stream of its forward node."""
        main = torch.cuda.current_stream()
        streams = _sub_streams(masked_kspace.device, ns)
        ks = masked_kspace.chunk(ns)
        rs = ref.chunk(ns) if ref is not None else [None] * len(ks)
        outs = []
        for st, k, r in zip(streams, ks, rs):
            st.wait_stream(main)
            mask.record_stream(st)          # a temporary of the caller, read on st until the backward has run
            with torch.cuda.stream(st):
                outs.append(self._forward_one(k, mask, r, num_low_frequencies))
        for st, o in zip(streams, outs):
            main.wait_stream(st)
            o.record_stream(main)           # allocated on st, read by the concatenation on main
        return torch.cat(outs, 0)

    def _forward_one(self, masked_kspace, mask, ref, num_low_frequencies):
        sens_maps = self.sens_net(masked_kspace, num_low_frequencies)
        kspace_pred = masked_kspace
        ref_normed = None
        if self.use_ref:
            ref = rss(ref)
            ref_normed = self.cascades[0].model.normalize_ref(ref) if len(self.cascades) else None
        for cascade in self.cascades:
            if self.checkpoint_cascades and torch.is_grad_enabled():
                from torch.utils.checkpoint import checkpoint
                kspace_pred = checkpoint(cascade, kspace_pred, masked_kspace, mask, sens_maps, ref, ref_normed,
                                         use_reentrant=False)
            else:
                kspace_pred = cascade(kspace_pred, masked_kspace, mask, sens_maps, ref, ref_normed)
        return ops.FftRss.apply(kspace_pred)


# Sub-batch streams of VarNet.forward (SAN_VARNET_STREAMS; default 1 = one chain on the current stream).  Measured on the
# B200 at bs 64 (profiles/r2t_*, r2u_*): 2 chains 390.9 ms/step against 369.1 eagerly (two chains double the launches and
# the host, at ~30 us per launch, no longer runs ahead of the GPU) and 386.3 against 371.3 even as ONE CUDA graph, 4 chains
# 420.3: the persistent conv CTAs (576 threads, 50 K registers per SM) leave the element-wise kernels of the other chain
# too little of each SM for the overlap to pay for the halved kernels.  Kept as an opt-in.
N_STREAMS = int(os.environ.get("SAN_VARNET_STREAMS", "1"))
MIN_SLICES_PER_STREAM = 4
_streams = {}


def _sub_streams(device, ns):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    have = _streams.setdefault(idx, [])
    while len(have) < ns:
        have.append(torch.cuda.Stream(device=idx))
    return have[:ns]
