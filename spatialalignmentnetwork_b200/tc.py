"""Fused convolution blocks on the tcgen05 path (``csrc/conv_tc.cu``).

The unit of the U-Nets is not "conv, then InstanceNorm, then LeakyReLU" but
``conv(act(norm(raw_1)) ++ act(norm(raw_2)) ...)``: the normalisation + activation (+ 2x2 average
pooling, pixel shuffle of the transposed conv, channel concat of the skip connection; reference
varnet.py:98,116,139-146,176-181) of the *producing* layers is applied while the operand of the
*consuming* convolution is staged as BF16 hi/lo tiles, so normalised / activated / concatenated /
pooled tensors are never written to HBM.  What crosses an autograd edge is always a raw fp32 conv
output; its per-plane statistics ride along in a ``Raw`` handle and their gradient is folded in
analytically (the InstanceNorm backward is linear in the incoming gradient, so every consumer
adds its own contribution).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import call, lib

MODE_DIRECT, MODE_POOL, MODE_D2S, MODE_UP = 0, 1, 2, 3
_IN_EPS = 1e-5


class Raw:
    """A raw fp32 NCHW tensor + how it is to be read by a consumer: ``norm`` in {None, 'in'}
    (InstanceNorm2d, biased variance, eps 1e-5), LeakyReLU ``slope`` (1 = none), and for ``d2s`` the
    tensor is the [N, 4C, h, w] output of the 1x1 form of ConvTranspose2d(2, stride 2), normalised
    over all four sub-planes of a channel."""

    def __init__(self, y, norm=None, slope=1.0, d2s=False):
        assert y.dtype == torch.float32 and y.dim() == 4
        self.y = y if y.is_contiguous() else y.contiguous()
        self.norm, self.slope, self.d2s = norm, float(slope), bool(d2s)
        self._coef = None

    @property
    def channels(self):
        return self.y.shape[1] // 4 if self.d2s else self.y.shape[1]

    def planes(self):
        N, C, H, W = self.y.shape
        return (N * C // 4, 4 * H * W) if self.d2s else (N * C, H * W)

    def coef(self):
        """[4, planes] = mean, m2, a (= rstd), b (= 0): computed once per raw tensor, shared by
        all its consumers (skip connection + pooled path)."""
        if self.norm is None:
            return None
        if self._coef is None:
            planes, P = self.planes()
            st = torch.empty(4, planes, dtype=torch.float32, device=self.y.device)
            yd = self.y.detach()
            call("plane_stats", yd, st[0], st[1], planes, P)
            call("in_finalize_fwd", st[0], st[1], st[2], st[3], planes, P, _IN_EPS)
            self._coef = st
        return self._coef


def _staged_act(N, H, W, C, device):
    return torch.empty(lib().san_tc_staged_act_elems(N, H, W, C), dtype=torch.bfloat16, device=device)


def _stage(xs, N, H, W, Cpad, srcs):
    """srcs: list of (y, coef | None, slope, C, mode), at most 3."""
    assert 1 <= len(srcs) <= 3
    flat = []
    for y, st, slope, C, mode in srcs:
        flat += [y, st[0] if st is not None else None, st[2] if st is not None else None, None, slope, C, mode]
    for _ in range(3 - len(srcs)):
        flat += [None, None, None, None, 1.0, 0, 0]
    call("tc_stage_act", xs, N, H, W, Cpad, *flat)


def _stage_weights(w, dgrad):
    Cout, Cin, K, _ = w.shape
    n = lib().san_tc_staged_weight_elems(Cin if dgrad else Cout, Cout if dgrad else Cin, K)
    ws = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    call("tc_stage_weights", w, ws, Cout, Cin, K, int(dgrad))
    return ws


def _pad16(c):
    return (c + 15) // 16 * 16


class _FusedConv(Function):
    """y = conv2d(concat_k act_k(norm_k(resample_k(raw_k))), w) + bias on the tcgen05 kernels."""

    @staticmethod
    def forward(ctx, w, bias, spec, *ys):
        # spec: (K, [(norm, slope, d2s, mode, coef)] per source)
        K, metas = spec
        w = w.contiguous()
        Cout, Cin, Kw, _ = w.shape
        assert Kw == K
        y0, m0 = ys[0], metas[0]
        N = y0.shape[0]
        if m0[3] == MODE_POOL:
            H, W = y0.shape[2] // 2, y0.shape[3] // 2
        elif m0[3] in (MODE_D2S, MODE_UP):
            H, W = y0.shape[2] * 2, y0.shape[3] * 2
        else:
            H, W = y0.shape[2], y0.shape[3]
        srcs, ctot = [], 0
        for y, (norm, slope, d2s, mode, coef) in zip(ys, metas):
            C = y.shape[1] // 4 if d2s else y.shape[1]
            srcs.append((y, coef, slope, C, mode))
            ctot += C
        assert ctot == Cin, (ctot, Cin)
        Cpad = _pad16(Cin)
        xs = _staged_act(N, H, W, Cin, w.device)
        _stage(xs, N, H, W, Cpad, srcs)
        ws = _stage_weights(w, False)
        out = torch.empty(N, Cout, H, W, dtype=torch.float32, device=w.device)
        call("tc_conv", xs, ws, bias, out, N, H, W, Cin, Cout, K, 0)
        # The staged operand is NOT kept for the backward (it is as large as the fp32 activation and
        # the padded channels make it larger): the weight gradient re-stages it from the raw tensors,
        # which autograd holds anyway for the normalisation backward.
        ctx.save_for_backward(w, *ys, *[m[4] for m in metas if m[4] is not None])
        ctx.meta = (K, [(m[0], m[1], m[2], m[3], m[4] is not None) for m in metas], (N, H, W), bias is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        K, metas, (N, H, W), has_bias = ctx.meta
        saved = ctx.saved_tensors
        w = saved[0]
        ns = len(metas)
        ys = saved[1:1 + ns]
        coefs = list(saved[1 + ns:])
        Cout, Cin = w.shape[0], w.shape[1]
        gy = gy if gy.is_contiguous() else gy.contiguous()
        dev = w.device
        # dY staged once as BF16 hi/lo: the operand of both the data- and the weight-gradient GEMMs
        gys = _staged_act(N, H, W, Cout, dev)
        _stage(gys, N, H, W, _pad16(Cout), [(gy, None, 1.0, Cout, MODE_DIRECT)])
        # ---- weight gradient: tcgen05 GEMM over the pixel dimension on the staged dY and the staged input
        dw = db = None
        if ctx.needs_input_grad[0] or (has_bias and ctx.needs_input_grad[1]):
            dw = torch.empty_like(w)
            db = torch.empty(Cout, dtype=torch.float32, device=dev) if has_bias else None
            xs = _staged_act(N, H, W, Cin, dev)          # re-stage the forward operand
            ci, srcs = 0, []
            for y, (norm, slope, d2s, mode, has_coef) in zip(ys, metas):
                srcs.append((y, coefs[ci] if has_coef else None, slope, y.shape[1] // 4 if d2s else y.shape[1], mode))
                ci += int(has_coef)
            _stage(xs, N, H, W, _pad16(Cin), srcs)
            if lib().san_tc_wgrad_supported(H, W, Cin, Cout, K):
                call("tc_wgrad", gys, xs, dw, db, gy if has_bias else None, N, H, W, Cin, Cout, K)
            else:   # tiny images (W < 16): fp32 CUDA-core kernel on the un-staged operand
                x32 = torch.empty(N, Cin, H, W, dtype=torch.float32, device=dev)
                call("tc_unstage_act", xs, x32, N, Cin, H, W)
                call("conv2d_wgrad", x32, gy, dw, db, N, Cin, H, W, Cout, K, 0, 0)
                del x32
            del xs
        # ---- data gradient: the same tcgen05 conv on the staged dY with the flipped filter
        grads = [None] * ns
        if any(ctx.needs_input_grad[3 + k] for k in range(ns)):
            wsd = _stage_weights(w, True)
            dx = torch.empty(N, Cin, H, W, dtype=torch.float32, device=dev)
            call("tc_conv", gys, wsd, None, dx, N, H, W, Cout, Cin, K, 0)
            del gys
            c0 = 0
            for k, (norm, slope, d2s, mode, has_coef) in enumerate(metas):
                y = ys[k]
                C = y.shape[1] // 4 if d2s else y.shape[1]
                if ctx.needs_input_grad[3 + k]:
                    g = dx[:, c0:c0 + C]
                    g = g if g.is_contiguous() else g.contiguous()
                    # undo the resampling: gradient at the resolution / layout of the raw tensor
                    if mode == MODE_POOL:
                        gs = torch.empty_like(y)
                        call("up2", g, gs, N * C, H, W, 0.25)
                    elif mode == MODE_D2S:
                        gs = torch.empty_like(y)
                        call("space_to_depth2", g, gs, N, C, H // 2, W // 2)
                    elif mode == MODE_UP:
                        gs = torch.empty_like(y)
                        call("pool2", g, gs, N * C, H, W, 1.0)
                    else:
                        gs = g
                    if has_coef:
                        st = coefs.pop(0)
                        planes = st.shape[1]
                        P = y.numel() // planes
                        wk = torch.empty(5, planes, dtype=torch.float32, device=dev)
                        call("act_bwd_reduce", gs, y, st[0], st[2], None, st[2], slope, wk[0], wk[1], planes, P)
                        call("in_finalize_bwd", wk[0], wk[1], st[2], wk[2], wk[3], wk[4], planes, P)
                        dy = torch.empty_like(y)
                        call("act_bwd_apply", gs, y, st[0], st[2], None, slope, wk[2], wk[3], wk[4], dy, planes, P)
                        grads[k] = dy
                    else:
                        assert slope == 1.0
                        grads[k] = gs
                elif has_coef:
                    coefs.pop(0)
                c0 += C
        return (dw, db, None, *grads)


def fused_conv(sources, weight, bias=None, modes=None):
    """sources: list of ``Raw``; modes: per-source resampling mode (default direct; a ``d2s`` source
    is read through the pixel shuffle).  Returns the raw fp32 conv output tensor."""
    modes = modes or [MODE_D2S if s.d2s else MODE_DIRECT for s in sources]
    metas = [(s.norm, s.slope, s.d2s, m, s.coef()) for s, m in zip(sources, modes)]
    K = weight.shape[-1]
    return _FusedConv.apply(weight, bias, (K, metas), *[s.y for s in sources])
