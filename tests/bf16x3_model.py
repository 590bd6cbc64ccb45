"""CPU error model of the tcgen05 convolution path (test infrastructure).

The CUDA convolutions split both operands into BF16 hi + lo halves and accumulate ``hi*hi + lo*hi + hi*lo`` in
fp32 (csrc/conv_tc.cu, csrc/wgrad_tc.cu), forward, data gradient and weight gradient alike.  ``patched()``
replaces ``torch.nn.functional.conv2d`` by an autograd function with exactly that rounding, so running the CPU
oracle under it predicts how far a *correct* BF16x3 implementation lands from the reference's fp32 results on a
given fixture - the principled way to set the bars of whole-network GPU tests on ill-conditioned tiny
networks (BatchNorm over a few dozen values amplifies operand rounding; NetG of the Mixed-step fixture has an
input-to-output error gain of 18-28x)."""
import contextlib

import torch
import torch.nn.functional as F

_conv2d = F.conv2d


def _split(t):
    hi = t.to(torch.bfloat16).to(t.dtype)
    return hi, (t - hi).to(torch.bfloat16).to(t.dtype)


def _x3(f, a, b):
    ah, al = _split(a)
    bh, bl = _split(b)
    return (f(ah.double(), bh.double()) + f(al.double(), bh.double()) + f(ah.double(), bl.double())).to(a.dtype)


class _ConvX3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.sp = (stride, padding)
        return _x3(lambda a, b: _conv2d(a, b, None, stride, padding), x, w)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding = ctx.sp
        gx = _x3(lambda a, b: torch.nn.grad.conv2d_input(x.shape, b, a, stride=stride, padding=padding), gy, w)
        gw = _x3(lambda a, b: torch.nn.grad.conv2d_weight(b, w.shape, a, stride=stride, padding=padding), gy, x)
        return gx, gw, None, None


def conv2d_bf16x3(x, w, b=None, stride=1, padding=0, *args, **kw):
    assert not args and not kw, "only the plain conv2d form the oracle uses"
    y = _ConvX3.apply(x, w, stride, padding)
    return y if b is None else y + b.view(1, -1, 1, 1)


@contextlib.contextmanager
def patched():
    F.conv2d = conv2d_bf16x3
    try:
        yield
    finally:
        F.conv2d = _conv2d
