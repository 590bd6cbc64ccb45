"""Oracle restatement of reference ``varnet.py`` (test infrastructure).

Functional form: every function takes a flat ``sd`` (``state_dict``-style
mapping name -> tensor, reference key names) and a key ``prefix``.
"""
import math

import torch
import torch.nn.functional as F

from .signal import fft2, ifft2, rss


def _inorm(x, eps=1e-5):
    """nn.InstanceNorm2d(affine=False): biased variance, eps inside the sqrt
    (varnet.py:141,144,180,235)."""
    # Same library call as the reference module: besides the forward values it
    # fixes the *backward* formula  dx = rstd*(g - mean(g) - xhat*mean(g*xhat)).
    # Autograd through an explicit (x-mean)/sqrt(var+eps) is measurably worse in
    # fp32 (3e-3 rel-L2 on the k-space gradient of tests/golden/varnet_s vs 3e-7).
    return F.instance_norm(x, eps=eps)


def conv_block(sd, p, x):
    """ConvBlock varnet.py:122-156: 2x [conv3x3 no-bias -> IN -> LeakyReLU(0.2)]."""
    x = F.conv2d(x, sd[p + "layers.0.weight"], padding=1)
    x = F.leaky_relu(_inorm(x), 0.2)
    x = F.conv2d(x, sd[p + "layers.3.weight"], padding=1)
    x = F.leaky_relu(_inorm(x), 0.2)
    return x


def transpose_conv_block(sd, p, x):
    """TransposeConvBlock varnet.py:159-192: convT 2x2 s2 no-bias -> IN -> LeakyReLU(0.2)."""
    x = F.conv_transpose2d(x, sd[p + "layers.0.weight"], stride=2)
    return F.leaky_relu(_inorm(x), 0.2)


def unet(sd, p, image, num_pool_layers):
    """Unet.forward varnet.py:82-119."""
    assert not torch.is_complex(image)
    stack = []
    out = image
    for i in range(num_pool_layers):
        out = conv_block(sd, f"{p}down_sample_layers.{i}.", out)
        stack.append(out)
        out = F.avg_pool2d(out, kernel_size=2, stride=2)
    out = conv_block(sd, p + "conv.", out)
    for i in range(num_pool_layers):
        skip = stack.pop()
        out = transpose_conv_block(sd, f"{p}up_transpose_conv.{i}.", out)
        pad = [0, 0, 0, 0]
        if out.shape[-1] != skip.shape[-1]:
            pad[1] = 1
        if out.shape[-2] != skip.shape[-2]:
            pad[3] = 1
        if sum(pad) != 0:
            out = F.pad(out, pad, "reflect")
        out = torch.cat([out, skip], dim=1)  # up-sampled first (varnet.py:116)
        if i < num_pool_layers - 1:
            out = conv_block(sd, f"{p}up_conv.{i}.", out)
        else:
            out = conv_block(sd, f"{p}up_conv.{i}.0.", out)
            out = F.conv2d(out, sd[f"{p}up_conv.{i}.1.weight"], sd[f"{p}up_conv.{i}.1.bias"])
    return out


def _pad16(x):
    """NormUnet.pad varnet.py:275-289."""
    _, _, h, w = x.shape
    w_mult = ((w - 1) | 15) + 1
    h_mult = ((h - 1) | 15) + 1
    w_pad = [math.floor((w_mult - w) / 2), math.ceil((w_mult - w) / 2)]
    h_pad = [math.floor((h_mult - h) / 2), math.ceil((h_mult - h) / 2)]
    return F.pad(x, w_pad + h_pad), (h_pad, w_pad, h_mult, w_mult)


def norm_unet(sd, p, x, ref, num_pools, use_ref):
    """NormUnet.forward varnet.py:301-332."""
    assert x.dim() == 4 and torch.is_complex(x)
    x = torch.cat([x.real, x.imag], dim=1)              # :246-248
    b, c, h, w = x.shape
    xg = x.reshape(b, 2, c // 2 * h * w)                # :257-268 group norm
    mean = xg.mean(dim=2).view(b, 2, 1, 1)
    std = xg.std(dim=2).view(b, 2, 1, 1)                # unbiased
    x = (x - mean) / (std + 1e-6)
    x, (h_pad, w_pad, h_mult, w_mult) = _pad16(x)
    if use_ref:
        assert not torch.is_complex(ref)
        r, _ = _pad16(_inorm(ref))                      # :315-318
        x = torch.cat([x, r], dim=1)
    else:
        assert ref is None
    x = unet(sd, p + "unet.", x, num_pools)
    x = x[..., h_pad[0]: h_mult - h_pad[1], w_pad[0]: w_mult - w_pad[1]]
    x = x * std + mean                                  # :270-273
    c2 = x.shape[1] // 2
    return torch.complex(x[:, :c2], x[:, c2:])


def sens_model(sd, p, masked_kspace, num_low_frequencies, num_pools):
    """SensitivityModel.forward varnet.py:389-420."""
    W = masked_kspace.shape[-1]
    acs = torch.ones(W)
    acs[num_low_frequencies:] = 0
    acs = torch.roll(acs, -num_low_frequencies // 2)    # python: (-n)//2
    acs = acs[None, None, None, :].to(masked_kspace)
    images = ifft2(acs * masked_kspace)
    N, C, H, Wd = images.shape
    s = norm_unet(sd, p + "norm_unet.", images.reshape(N * C, 1, H, Wd), None, num_pools, False)
    s = s.reshape(N, C, H, Wd)
    return s / (rss(s) + 1e-6)


def varnet_block(sd, p, k, k0, mask, sens, ref, num_pools, use_ref):
    """VarNetBlock.forward varnet.py:514-530."""
    x = (ifft2(k) * sens.conj()).sum(dim=1, keepdim=True)          # sens_reduce :511-512
    x = norm_unet(sd, p + "model.", x, ref if use_ref else None, num_pools, use_ref)
    m = fft2(x * sens)                                             # sens_expand :508-509
    zero = torch.zeros(1, 1, 1, 1).to(k)
    soft_dc = torch.where(mask, k - k0, zero) * sd[p + "dc_weight"]
    return k - soft_dc - m


def varnet(sd, p, masked_kspace, mask, ref, num_low_frequencies,
           num_cascades, sens_pools=4, pools=4, use_ref=True):
    """VarNet.forward varnet.py:465-486."""
    sens = sens_model(sd, p + "sens_net.", masked_kspace, num_low_frequencies, sens_pools)
    k = masked_kspace.clone()
    if use_ref:
        ref = rss(ref)
    for i in range(num_cascades):
        k = varnet_block(sd, f"{p}cascades.{i}.", k, masked_kspace, mask, sens, ref, pools, use_ref)
    return rss(ifft2(k))
