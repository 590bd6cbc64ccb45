"""CPU stand-ins for the C-ABI ops the fused network walkers call (test infrastructure).

``tests/test_cpu_gan_host.py`` monkeypatches them into ``spatialalignmentnetwork_b200.{tc, ops}`` so that the
HOST logic of the walkers (which tensors are concatenated / summed, which BatchNorm slice normalises which
source, weight re-orderings, buffer updates) can be checked against the reference's golden vectors without a
GPU.  Each stand-in states the documented semantics of the op it replaces (include/san_b200.h,
spatialalignmentnetwork_b200/tc.py) with plain torch CPU calls; none of this is reachable from the package."""
import torch
import torch.nn.functional as F

MODE_DIRECT, MODE_POOL, MODE_D2S, MODE_UP = 0, 1, 2, 3


def _term(r, mode):
    """act(norm(resample(raw))) of one ``tc.Raw`` term, in the reference's order of operations."""
    x = r.y
    if r.d2s:                                   # ConvTranspose2d(2, 2) as 1x1 conv + pixel shuffle, then InstanceNorm
        x = F.pixel_shuffle(x, 2)
    if mode == MODE_UP:                         # nn.Upsample(nearest x2) BEFORE the normalisation (unet.py:128-133)
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
    if mode == MODE_POOL:                       # avg_pool2d of the ACTIVATED tensor (varnet.py:98)
        x = F.avg_pool2d(x, 2)
    return x


def fused_conv(sources, weight, bias=None, modes=None):
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


def install(monkeypatch):
    """Route the walkers' op calls to the stand-ins above."""
    from spatialalignmentnetwork_b200 import ops, tc
    monkeypatch.setattr(tc, "fused_conv", fused_conv)
    monkeypatch.setattr(ops, "add", lambda a, b: a + b)
    monkeypatch.setattr(ops, "BatchNormLReLU", _Apply(_bn_lrelu))
    monkeypatch.setattr(ops, "SpaceToDepth2", _Apply(lambda x: F.pixel_unshuffle(x, 2)))
    monkeypatch.setattr(ops, "AvgPool2", _Apply(lambda x: F.avg_pool2d(x, 2)))
    monkeypatch.setattr(ops, "SpectralNormWeight", _Apply(_sn_weight))
    monkeypatch.setattr(ops, "PairLoss", _Apply(_pair_loss))
