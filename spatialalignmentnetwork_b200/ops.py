"""``torch.autograd.Function`` wrappers over the C ABI (``include/san_b200.h``).

Every op here launches hand-written sm_100a kernels from ``libsan_b200.so`` on the
current CUDA stream; PyTorch only owns the memory and chains the backward passes.
There is no CPU path: inputs must be contiguous CUDA tensors (fp32 / complex64).
"""
import math

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import call


def _c(x):
    if x.is_conj() or x.is_neg():       # lazy .conj() / negative-bit views share storage: the kernels read raw memory
        x = x.resolve_conj().resolve_neg()
    return x if x.is_contiguous() else x.contiguous()


def _f32(x, name):
    assert x.dtype == torch.float32, f"{name}: expected float32, got {x.dtype}"
    return _c(x)


def _c64(x, name):
    assert x.dtype == torch.complex64, f"{name}: expected complex64, got {x.dtype}"
    return _c(x)


# --------------------------------------------------------------------------- FFT / DC
def _fft_plain(x, inverse, in_planar=False, out_planar=False, mask_in=None, mask_out=None):
    """x: complex64 [B,H,W] (or planar float [B,2,H,W]) -> same layouts."""
    if in_planar:
        B, _, H, W = x.shape
    else:
        B, H, W = x.shape
    if out_planar:
        out = torch.empty(B, 2, H, W, dtype=torch.float32, device=x.device)
        tmp = torch.empty(B, H, W, dtype=torch.complex64, device=x.device)
    else:
        out = torch.empty(B, H, W, dtype=torch.complex64, device=x.device)
        tmp = out
    call("fft2", x, int(in_planar), mask_in, out, int(out_planar), mask_out, tmp, B, H, W, int(inverse))
    return out


class Fft2(Function):
    """fft2 / ifft2 (norm='ortho') over the last two dims of a complex64 tensor
    (reference signal_utils.py:4-12); adjoint = the opposite transform."""

    @staticmethod
    def forward(ctx, x, inverse):
        x = _c64(x, "fft2")
        ctx.inverse = inverse
        shp = x.shape
        return _fft_plain(x.reshape(-1, shp[-2], shp[-1]), inverse).reshape(shp)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _c64(g, "fft2.backward")
        shp = g.shape
        return _fft_plain(g.reshape(-1, shp[-2], shp[-1]), not ctx.inverse).reshape(shp), None


class IfftMaskedPlanar(Function):
    """planar(ifft2(colmask * k)) for the sensitivity estimator (reference varnet.py:395-407):
    k complex64 [N,C,H,W], colmask float [W] -> float [N*C, 2, H, W]."""

    @staticmethod
    def forward(ctx, k, colmask):
        k = _c64(k, "ifft_masked")
        N, C, H, W = k.shape
        ctx.save_for_backward(colmask)
        ctx.shape = k.shape
        return _fft_plain(k.reshape(N * C, H, W), True, out_planar=True, mask_in=colmask)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (colmask,) = ctx.saved_tensors
        g = _f32(g, "ifft_masked.backward")
        return _fft_plain(g, False, in_planar=True, mask_out=colmask).reshape(ctx.shape), None


class FftReduce(Function):
    """sens_reduce (reference varnet.py:511-512): planar(sum_c ifft2(k) * conj(S)) -> [N,2,H,W]."""

    @staticmethod
    def forward(ctx, k, sens):
        k, sens = _c64(k, "fft_reduce"), _c64(sens, "fft_reduce")
        N, C, H, W = k.shape
        x = torch.empty(N, 2, H, W, dtype=torch.float32, device=k.device)
        need_u = ctx.needs_input_grad[1]
        u = torch.empty_like(k) if need_u else None
        tmp = torch.empty_like(k)
        call("fft_reduce", k, sens, x, u, tmp, N, C, H, W, 1, 1.0)
        ctx.save_for_backward(sens, u)
        return x

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        sens, u = ctx.saved_tensors
        g = _f32(g, "fft_reduce.backward")
        N, C, H, W = sens.shape
        dk = dS = None
        if ctx.needs_input_grad[0]:
            dk = torch.empty_like(sens)
            call("fft_expand_dc", g, sens, None, None, None, None, dk, dk, N, C, H, W, 0)
        if ctx.needs_input_grad[1]:
            dS = torch.empty_like(sens)
            call("cmul_conj_planar", u, g, dS, N, C, H * W, 1.0)
        return dk, dS


class FftExpandDC(Function):
    """k' = k - where(mask, k - k0, 0) * dc_weight - fft2(x * S)
    (reference varnet.py:508-509, 525-530); x planar [N,2,H,W]."""

    @staticmethod
    def forward(ctx, x, sens, k, k0, mask, dc_weight):
        x = _f32(x, "fft_expand_dc")
        sens, k, k0 = _c64(sens, "fft_expand_dc"), _c64(k, "fft_expand_dc"), _c64(k0, "fft_expand_dc")
        assert mask.dtype == torch.bool and mask.dim() == 1 and mask.shape[0] == k.shape[-1]
        mask = _c(mask)
        N, C, H, W = k.shape
        out = torch.empty_like(k)
        tmp = torch.empty_like(k)
        call("fft_expand_dc", x, sens, k, k0, mask, dc_weight, out, tmp, N, C, H, W, 0)
        ctx.save_for_backward(x, sens, k, k0, mask, dc_weight)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        x, sens, k, k0, mask, dc_weight = ctx.saved_tensors
        G = _c64(G, "fft_expand_dc.backward")
        N, C, H, W = k.shape
        dx = dS = dk = dk0 = dw = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dx = torch.empty_like(x)
            u = torch.empty_like(k) if ctx.needs_input_grad[1] else None
            tmp = torch.empty_like(k)
            call("fft_reduce", G, sens, dx, u, tmp, N, C, H, W, 1, -1.0)   # u = -ifft2(G)
            if u is not None:
                dS = torch.empty_like(k)
                call("cmul_conj_planar", u, x, dS, N, C, H * W, 1.0)
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3] or ctx.needs_input_grad[5]:
            dk = torch.empty_like(k) if ctx.needs_input_grad[2] else None
            dk0 = torch.empty_like(k) if ctx.needs_input_grad[3] else None
            dw = torch.empty_like(dc_weight)
            scratch = torch.empty(1, dtype=torch.float64, device=k.device)
            call("dc_bwd", G, k, k0, mask, dc_weight, dk, dk0, dw, scratch, N * C * H, W)
        return dx, dS, dk, dk0, None, dw


class FftRss(Function):
    """rss(ifft2(k)) (reference varnet.py:486): complex64 [N,C,H,W] -> float [N,1,H,W]."""

    @staticmethod
    def forward(ctx, k):
        k = _c64(k, "fft_rss")
        N, C, H, W = k.shape
        out = torch.empty(N, 1, H, W, dtype=torch.float32, device=k.device)
        u = torch.empty_like(k)
        tmp = torch.empty_like(k)
        call("fft_rss", k, out, u, tmp, N, C, H, W, 1)
        ctx.save_for_backward(u, out)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        u, r = ctx.saved_tensors
        g = _f32(g, "fft_rss.backward")
        N, C, H, W = u.shape
        du = torch.empty_like(u)
        call("rss_bwd", g, u, r, du, N, C, H * W, 1)
        dk = _fft_plain(du.reshape(N * C, H, W), False).reshape(u.shape)
        return dk


class Rss(Function):
    """signal_utils.rss (reference signal_utils.py:24-26): L2 norm over dim 1, keepdim."""

    @staticmethod
    def forward(ctx, x):
        assert x.dim() == 4
        x = _c(x)
        is_c = x.dtype == torch.complex64
        assert is_c or x.dtype == torch.float32
        N, C, H, W = x.shape
        out = torch.empty(N, 1, H, W, dtype=torch.float32, device=x.device)
        call("rss_fwd", x, out, N, C, H * W, int(is_c))
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, r = ctx.saved_tensors
        g = _f32(g, "rss.backward")
        N, C, H, W = x.shape
        dx = torch.empty_like(x)
        call("rss_bwd", g, x, r, dx, N, C, H * W, int(x.dtype == torch.complex64))
        return dx


class SensNormalize(Function):
    """S = s / (rss(s) + 1e-6) (reference varnet.py:419); s planar [N*C,2,H,W] -> complex64 [N,C,H,W]."""

    @staticmethod
    def forward(ctx, s, N, C):
        s = _f32(s, "sens_normalize")
        _, _, H, W = s.shape
        S = torch.empty(N, C, H, W, dtype=torch.complex64, device=s.device)
        call("sens_normalize_fwd", s, S, N, C, H * W, 1e-6)
        ctx.save_for_backward(s)
        ctx.nc = (N, C)
        return S

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        (s,) = ctx.saved_tensors
        N, C = ctx.nc
        G = _c64(G, "sens_normalize.backward")
        ds = torch.empty_like(s)
        call("sens_normalize_bwd", G, s, ds, N, C, s.shape[2] * s.shape[3], 1e-6)
        return ds, None, None


# --------------------------------------------------------------------------- convolutions
def _pack(w, dgrad):
    Cout, Cin, K, _ = w.shape
    p = torch.empty(w.numel(), dtype=torch.float32, device=w.device)
    call("conv_pack_weights", w, p, Cout, Cin, K, int(dgrad))
    return p


class Conv2d(Function):
    """conv2d stride 1, padding K/2, K in {1,3}, NCHW fp32 (reference varnet.py:140,143,78;
    unet.py:123,131,138,186; cross.py:15)."""

    @staticmethod
    def forward(ctx, x, w, bias):
        x, w = _f32(x, "conv2d"), _f32(w, "conv2d")
        N, Cin, H, W = x.shape
        Cout, Cin2, K, K2 = w.shape
        assert Cin == Cin2 and K == K2 and K in (1, 3), (x.shape, w.shape)
        y = torch.empty(N, Cout, H, W, dtype=torch.float32, device=x.device)
        call("conv2d_fwd", x, _pack(w, False), bias, y, N, Cin, H, W, Cout, K, 0, 0)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32(dy, "conv2d.backward")
        N, Cin, H, W = x.shape
        Cout, _, K, _ = w.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("conv2d_fwd", dy, _pack(w, True), None, dx, N, Cout, H, W, Cin, K, 0, 0)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.empty_like(w)
            db = torch.empty(Cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            call("conv2d_wgrad", x, dy, dw, db, N, Cin, H, W, Cout, K, 0, 0)
        return dx, dw, db


class DepthToSpace2(Function):
    """[N, Co*4, H, W] (channel = co*4 + a*2 + b) -> [N, Co, 2H, 2W]."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "depth_to_space")
        N, C4, H, W = x.shape
        y = torch.empty(N, C4 // 4, 2 * H, 2 * W, dtype=torch.float32, device=x.device)
        call("depth_to_space2", x, y, N, C4 // 4, H, W)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _f32(g, "depth_to_space.backward")
        N, Co, H2, W2 = g.shape
        dx = torch.empty(N, Co * 4, H2 // 2, W2 // 2, dtype=torch.float32, device=g.device)
        call("space_to_depth2", g, dx, N, Co, H2 // 2, W2 // 2)
        return dx


def conv_transpose2x2(x, w):
    """ConvTranspose2d(kernel 2, stride 2, no bias) (reference varnet.py:176-179); w [Cin,Cout,2,2]."""
    Cin, Cout = w.shape[0], w.shape[1]
    w1 = w.permute(1, 2, 3, 0).reshape(Cout * 4, Cin, 1, 1)
    return DepthToSpace2.apply(Conv2d.apply(x, w1, None))


# --------------------------------------------------------------------------- norm / act
def _planes(y):
    N, C, H, W = y.shape
    return N * C, H * W


class InstanceNormLReLU(Function):
    """InstanceNorm2d(affine=False, biased var) + LeakyReLU (reference varnet.py:141-145,180-181);
    slope=1 gives the bare InstanceNorm of varnet.py:235."""

    @staticmethod
    def forward(ctx, y, slope, eps):
        y = _f32(y, "instance_norm")
        planes, P = _planes(y)
        st = torch.empty(4, planes, dtype=torch.float32, device=y.device)   # mean, m2, a (= rstd), b (= 0)
        call("plane_stats", y, st[0], st[1], planes, P)
        call("in_finalize_fwd", st[0], st[1], st[2], st[3], planes, P, eps)
        out = torch.empty_like(y)
        call("affine_act_fwd", y, st[0], st[2], None, slope, out, planes, P)
        ctx.save_for_backward(y, st)
        ctx.slope = slope
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, st = ctx.saved_tensors
        g = _f32(g, "instance_norm.backward")
        planes, P = _planes(y)
        mu, a = st[0], st[2]
        w = torch.empty(5, planes, dtype=torch.float32, device=y.device)
        call("act_bwd_reduce", g, y, mu, a, None, a, ctx.slope, w[0], w[1], planes, P)
        call("in_finalize_bwd", w[0], w[1], a, w[2], w[3], w[4], planes, P)
        dy = torch.empty_like(y)
        call("act_bwd_apply", g, y, mu, a, None, ctx.slope, w[2], w[3], w[4], dy, planes, P)
        return dy, None, None


class BatchNormLReLU(Function):
    """BatchNorm2d(affine, running stats) + LeakyReLU (reference unet.py:124-126)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, training, momentum, eps, slope):
        y = _f32(y, "batch_norm")
        N, C, H, W = y.shape
        planes, P = N * C, H * W
        st = torch.empty(6, planes, dtype=torch.float32, device=y.device)   # mean, m2, mu, a, b, sa
        if training:
            call("plane_stats", y, st[0], st[1], planes, P)
        call("bn_finalize_fwd", st[0], st[1], gamma, beta, running_mean, running_var, st[2], st[3], st[4], st[5],
             N, C, P, eps, momentum, int(training))
        out = torch.empty_like(y)
        call("affine_act_fwd", y, st[2], st[3], st[4], slope, out, planes, P)
        ctx.save_for_backward(y, st, gamma)
        ctx.cfg = (slope, bool(training))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, st, gamma = ctx.saved_tensors
        slope, training = ctx.cfg
        g = _f32(g, "batch_norm.backward")
        N, C, H, W = y.shape
        planes, P = N * C, H * W
        mu, a, b, sa = st[2], st[3], st[4], st[5]
        w = torch.empty(5, planes, dtype=torch.float32, device=y.device)
        call("act_bwd_reduce", g, y, mu, a, b, sa, slope, w[0], w[1], planes, P)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        call("bn_finalize_bwd", w[0], w[1], gamma, sa, w[2], w[3], w[4], dgamma, dbeta, N, C, P, int(training))
        dy = torch.empty_like(y)
        call("act_bwd_apply", g, y, mu, a, b, slope, w[2], w[3], w[4], dy, planes, P)
        return dy, dgamma, dbeta, None, None, None, None, None, None


class PlaneStats(Function):
    """Per-(n,c)-plane mean and centred sum of squares -> (mean [N,C], m2 [N,C])."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "plane_stats")
        N, C, H, W = x.shape
        st = torch.empty(2, N * C, dtype=torch.float32, device=x.device)
        call("plane_stats", x, st[0], st[1], N * C, H * W)
        ctx.save_for_backward(x, st)
        return st[0].view(N, C), st[1].view(N, C)

    @staticmethod
    @once_differentiable
    def backward(ctx, gmean, gm2):
        # d mean / dx = 1/P ; d m2 / dx = 2 (x - mean)
        x, st = ctx.saved_tensors
        N, C, H, W = x.shape
        P = H * W
        a = (2.0 * gm2).reshape(-1).contiguous()
        b = (gmean.reshape(-1) / P).contiguous()
        dx = torch.empty_like(x)
        call("affine_act_fwd", x, st[0], a, b, 1.0, dx, N * C, P)
        return dx


class PlaneAffine(Function):
    """out[n,c] = a[n,c] * (x[n,c] - mu[n,c]) + b[n,c] with per-plane scalars (mu, b optional),
    differentiable in all four; the norm / unnorm of NormUnet (reference varnet.py:257-273) in the
    centred form PyTorch's ``(x - mean) / std`` has."""

    @staticmethod
    def forward(ctx, x, mu, a, b):
        x = _f32(x, "plane_affine")
        N, C, H, W = x.shape
        a = _f32(a.reshape(-1), "plane_affine")
        mu = _f32(mu.reshape(-1), "plane_affine") if mu is not None else None
        b = _f32(b.reshape(-1), "plane_affine") if b is not None else None
        out = torch.empty_like(x)
        call("affine_act_fwd", x, mu, a, b, 1.0, out, N * C, H * W)
        ctx.save_for_backward(x, mu, a)
        ctx.has = (mu is not None, b is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, mu, a = ctx.saved_tensors
        g = _f32(g, "plane_affine.backward")
        N, C, H, W = x.shape
        planes, P = N * C, H * W
        s = torch.empty(2, planes, dtype=torch.float32, device=x.device)   # sum g, sum g*(x - mu)
        call("act_bwd_reduce", g, x, mu, a, None, None, 1.0, s[0], s[1], planes, P)
        dx = torch.empty_like(x)
        call("act_bwd_apply", g, x, mu, a, None, 1.0, a, None, None, dx, planes, P)
        dmu = (-a * s[0]).view(N, C) if ctx.has[0] else None
        db = s[0].view(N, C) if ctx.has[1] else None
        return dx, dmu, s[1].view(N, C), db


class AvgPool2(Function):
    """F.avg_pool2d(kernel 2, stride 2) (reference varnet.py:98, unet.py:137)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "avg_pool2")
        N, C, H, W = x.shape
        y = torch.empty(N, C, H // 2, W // 2, dtype=torch.float32, device=x.device)
        call("pool2", x, y, N * C, H, W, 0.25)
        ctx.shape = x.shape
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _f32(g, "avg_pool2.backward")
        N, C, H, W = ctx.shape
        assert H % 2 == 0 and W % 2 == 0, "avg_pool2 backward needs even sizes"
        dx = torch.empty(N, C, H, W, dtype=torch.float32, device=g.device)
        call("up2", g, dx, N * C, H // 2, W // 2, 0.25)
        return dx


class Upsample2(Function):
    """nn.Upsample(scale_factor=2, mode='nearest') (reference unet.py:130)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "upsample2")
        N, C, H, W = x.shape
        y = torch.empty(N, C, 2 * H, 2 * W, dtype=torch.float32, device=x.device)
        call("up2", x, y, N * C, H, W, 1.0)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _f32(g, "upsample2.backward")
        N, C, H2, W2 = g.shape
        dx = torch.empty(N, C, H2 // 2, W2 // 2, dtype=torch.float32, device=g.device)
        call("pool2", g, dx, N * C, H2, W2, 1.0)
        return dx


# --------------------------------------------------------------------------- alignment
class GridFromOffset(Function):
    """identity affine_grid(align_corners=False) + offset (reference cross.py:24-29);
    x = network output [N,2,H,W] -> grid [N,H,W,2]."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "grid_from_offset")
        N, two, H, W = x.shape
        assert two == 2
        grid = torch.empty(N, H, W, 2, dtype=torch.float32, device=x.device)
        call("grid_from_offset", x, grid, N, H, W)
        return grid

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _f32(g, "grid_from_offset.backward")
        N, H, W, _ = g.shape
        dx = torch.empty(N, 2, H, W, dtype=torch.float32, device=g.device)
        call("grid_to_nchw", g, dx, N, H, W)
        return dx


class Warp(Function):
    """F.grid_sample(bilinear, zeros, align_corners=False) (reference cross.py:32-38)."""

    @staticmethod
    def forward(ctx, img, grid):
        img, grid = _f32(img, "warp"), _f32(grid, "warp")
        N, C, H, W = img.shape
        N2, Ho, Wo, two = grid.shape
        assert N == N2 and two == 2
        out = torch.empty(N, C, Ho, Wo, dtype=torch.float32, device=img.device)
        call("warp_fwd", img, grid, out, N, C, H, W, Ho, Wo)
        ctx.save_for_backward(img, grid)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        img, grid = ctx.saved_tensors
        g = _f32(g, "warp.backward")
        N, C, H, W = img.shape
        _, Ho, Wo, _ = grid.shape
        dimg = torch.empty_like(img) if ctx.needs_input_grad[0] else None
        dgrid = torch.empty_like(grid) if ctx.needs_input_grad[1] else None
        if dimg is None and dgrid is None:
            return None, None
        call("warp_bwd", g, img, grid, dimg, dgrid, N, C, H, W, Ho, Wo)
        return dimg, dgrid


class GradientLoss(Function):
    """Displacement smoothness (reference model.py:21-28); s [N,H,W,2] with any strides."""

    @staticmethod
    def forward(ctx, s):
        assert s.shape[-1] == 2 and s.dim() == 4 and s.dtype == torch.float32
        N, H, W, _ = s.shape
        out = torch.empty((), dtype=torch.float32, device=s.device)
        scratch = torch.empty(2, dtype=torch.float64, device=s.device)
        st = s.stride()
        call("grad_loss_fwd", s.data_ptr(), st[0], st[1], st[2], st[3], N, H, W, out, scratch)
        ctx.save_for_backward(s)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        N, H, W, _ = s.shape
        ds = torch.empty(N, H, W, 2, dtype=torch.float32, device=s.device)
        st = s.stride()
        call("grad_loss_bwd", s.data_ptr(), st[0], st[1], st[2], st[3], N, H, W, _c(g), ds)
        return ds


# --------------------------------------------------------------------------- losses
class _WindowLoss(Function):
    NAME = None

    @classmethod
    def _fwd(cls, ctx, X, Y):
        X, Y = _f32(X, cls.NAME), _f32(Y, cls.NAME)
        N, C, H, W = X.shape
        assert C == 1 and X.shape == Y.shape, f"{cls.NAME}: single-channel images of equal shape"
        out = torch.empty((), dtype=torch.float32, device=X.device)
        scratch = torch.empty(1, dtype=torch.float64, device=X.device)
        call(cls.NAME + "_loss_fwd", X, Y, N, H, W, out, scratch)
        ctx.save_for_backward(X, Y)
        return out

    @classmethod
    def _bwd(cls, ctx, g):
        X, Y = ctx.saved_tensors
        N, C, H, W = X.shape
        dX = torch.empty_like(X) if ctx.needs_input_grad[0] else None
        dY = torch.empty_like(Y) if ctx.needs_input_grad[1] else None
        if dX is None and dY is None:
            return None, None
        call(cls.NAME + "_loss_bwd", X, Y, _c(g), N, H, W, dX, dY)
        return dX, dY


class SsimLoss(_WindowLoss):
    """1 - mean SSIM, 7x7 uniform window (reference ssimloss.py:11-40)."""
    NAME = "ssim"

    @staticmethod
    def forward(ctx, X, Y):
        return SsimLoss._fwd(ctx, X, Y)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return SsimLoss._bwd(ctx, g)


class LnccLoss(_WindowLoss):
    """-mean local normalised cross-correlation, 9x9 window (reference lnccloss.py:7-56)."""
    NAME = "lncc"

    @staticmethod
    def forward(ctx, X, Y):
        return LnccLoss._fwd(ctx, X, Y)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return LnccLoss._bwd(ctx, g)


class MiHist(Function):
    """Gaussian Parzen soft histograms (reference miloss.py:26-42): I, J [N,P] ->
    joint [N,64,64] = p_I p_J^T, mI [N,64], mJ [N,64] (row sums of p_I, p_J)."""

    @staticmethod
    def forward(ctx, I, J, bins, sigma, minv, maxv):
        I, J = _f32(I, "mi_hist"), _f32(J, "mi_hist")
        N, P = I.shape
        joint = torch.empty(N, bins, bins, dtype=torch.float32, device=I.device)
        mI = torch.empty(N, bins, dtype=torch.float32, device=I.device)
        mJ = torch.empty(N, bins, dtype=torch.float32, device=I.device)
        call("mi_hist_fwd", I, J, joint, mI, mJ, N, P, bins, sigma, minv, maxv)
        ctx.save_for_backward(I, J)
        ctx.cfg = (bins, sigma, minv, maxv)
        return joint, mI, mJ

    @staticmethod
    @once_differentiable
    def backward(ctx, gj, gi, gjj):
        I, J = ctx.saved_tensors
        bins, sigma, minv, maxv = ctx.cfg
        N, P = I.shape
        dI = torch.empty_like(I) if ctx.needs_input_grad[0] else None
        dJ = torch.empty_like(J) if ctx.needs_input_grad[1] else None
        if dI is None and dJ is None:
            return None, None, None, None, None, None
        call("mi_hist_bwd", I, J, _f32(gj, "mi"), _f32(gi, "mi"), _f32(gjj, "mi"), dI, dJ, N, P, bins, sigma, minv, maxv)
        return dI, dJ, None, None, None, None


class Filter2d(Function):
    """Single-channel KxK correlation with zero padding K/2 on every plane of x."""

    @staticmethod
    def forward(ctx, x, w):
        x, w = _f32(x, "filter2d"), _f32(w, "filter2d")
        K = w.shape[-1]
        N, C, H, W = x.shape
        y = torch.empty_like(x)
        call("filter2d", x, w, y, N * C, H, W, K)
        ctx.save_for_backward(w)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        g = _f32(g, "filter2d.backward")
        K = w.shape[-1]
        N, C, H, W = g.shape
        dx = torch.empty_like(g)
        call("filter2d", g, w.flip(-1, -2).contiguous(), dx, N * C, H, W, K)
        return dx, None


def add(x, y):
    """Residual add (reference unet.py:23) through the library's axpby kernel."""
    return _Add.apply(x, y)


class _Add(Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _f32(x, "add"), _f32(y, "add")
        out = torch.empty_like(x)
        call("axpby", x, y, out, 1.0, 1.0, x.numel())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


class SpaceToDepth2(Function):
    """[N, C, 2H, 2W] -> [N, C*4, H, W] (channel = c*4 + a*2 + b): turns the kernel-2 stride-2 convolution of
    gan.py:43-46 into a 1x1 convolution over 4C channels."""

    @staticmethod
    def forward(ctx, y):
        y = _f32(y, "space_to_depth")
        N, C, H2, W2 = y.shape
        assert H2 % 2 == 0 and W2 % 2 == 0, "space_to_depth needs even sizes"
        x = torch.empty(N, C * 4, H2 // 2, W2 // 2, dtype=torch.float32, device=y.device)
        call("space_to_depth2", y, x, N, C, H2 // 2, W2 // 2)
        return x

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _f32(g, "space_to_depth.backward")
        N, C4, H, W = g.shape
        dy = torch.empty(N, C4 // 4, 2 * H, 2 * W, dtype=torch.float32, device=g.device)
        call("depth_to_space2", g, dy, N, C4 // 4, H, W)
        return dy


# --------------------------------------------------------------------------- GAN branch
class SpectralNormWeight(Function):
    """``torch.nn.utils.spectral_norm`` (reference gan.py:24): W / sigma with sigma = u . (W v); in training
    mode one power iteration first updates the ``u`` / ``v`` buffers in place.  u, v are constants of the
    backward (torch detaches / clones them)."""

    @staticmethod
    def forward(ctx, w, u, v, power_iteration, eps):
        w = _f32(w, "spectral_norm")
        rows = w.shape[0]
        cols = w.numel() // rows
        assert u.numel() == rows and v.numel() == cols and u.is_contiguous() and v.is_contiguous()
        tmp = torch.empty(max(rows, cols), dtype=torch.float32, device=w.device)
        sigma = torch.empty(1, dtype=torch.float32, device=w.device)
        call("sn_sigma", w, u, v, tmp, sigma, rows, cols, eps, int(power_iteration))
        out = torch.empty_like(w)
        call("sn_scale", w, sigma, out, w.numel())
        ctx.save_for_backward(out, u.clone(), v.clone(), sigma)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        w_sn, u, v, sigma = ctx.saved_tensors
        g = _f32(g, "spectral_norm.backward")
        rows = w_sn.shape[0]
        cols = w_sn.numel() // rows
        dw = torch.empty_like(w_sn)
        scratch = torch.empty(1, dtype=torch.float64, device=g.device)
        call("sn_bwd", g, w_sn, u, v, sigma, scratch, dw, rows, cols)
        return dw, None, None, None, None


class PairLoss(Function):
    """mean_i f(x_i, y_i): mode 0 ``F.l1_loss`` (model.py:138-139), mode 1 ``clamp(sign*x, min=-1).mean()``
    and mode 2 ``(sign*x).mean()`` (the hinge / generator terms of ``loss_gan``, gan.py:131-137)."""

    @staticmethod
    def forward(ctx, x, y, mode, sign):
        x = _f32(x, "pair_loss")
        y = _f32(y, "pair_loss") if y is not None else None
        assert mode in (0, 1, 2) and (mode != 0 or (y is not None and y.shape == x.shape))
        out = torch.empty((), dtype=torch.float32, device=x.device)
        scratch = torch.empty(1, dtype=torch.float64, device=x.device)
        call("pair_loss_fwd", x, y, x.numel(), mode, float(sign), out, scratch)
        ctx.save_for_backward(x, y)
        ctx.cfg = (mode, float(sign))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        mode, sign = ctx.cfg
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dy = torch.empty_like(y) if (y is not None and ctx.needs_input_grad[1]) else None
        if dx is None and dy is None:
            return None, None, None, None
        if dx is None and mode != 0:
            return None, None, None, None
        call("pair_loss_bwd", x, y, _c(g), x.numel(), mode, sign, dx, dy)
        return dx, dy, None, None


def l1_loss(x, y):
    """``torch.nn.functional.l1_loss`` (mean reduction)."""
    return PairLoss.apply(x, y, 0, 1.0)


# --------------------------------------------------------------------------- metrics
def error_sums(a, b):
    """-> (sum (a-b)^2, sum |a-b|, sum a^2) as python floats (fp64 device reduction, one D2H of 24 bytes)."""
    a, b = _f32(a, "error_sums"), _f32(b, "error_sums")
    assert a.shape == b.shape
    out = torch.empty(3, dtype=torch.float64, device=a.device)
    call("error_sums", a, b, a.numel(), out)
    return tuple(out.tolist())


def mi_metric(gt, pred, bins=64, minVal=0.0, maxVal=1.0):
    """metrics.mi (reference metrics.py:54-68): mean over the batch of the plug-in mutual information of the
    hard ``bins x bins`` joint histogram."""
    gt, pred = _f32(gt, "mi_metric"), _f32(pred, "mi_metric")
    assert gt.shape == pred.shape and gt.dim() == 4, "wrong shape [batch, channel=1, rows, cols"
    N = gt.shape[0]
    out = torch.empty(N, dtype=torch.float64, device=gt.device)
    call("mi_metric", gt, pred, N, gt.numel() // N, bins, float(minVal), float(maxVal), out)
    vals = out.tolist()
    return sum(vals) / len(vals)
