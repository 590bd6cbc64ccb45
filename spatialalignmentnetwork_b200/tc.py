"""Fused convolution blocks on the tcgen05 path (``csrc/conv_tc.cu``, ``csrc/wgrad_tc.cu``).

The unit of the U-Nets is not "conv, then norm, then LeakyReLU" but
``conv(concat_k sum_j act(norm(raw_kj)))``: the normalisation + activation (+ 2x2 average pooling, pixel
shuffle of the transposed conv, nearest up-sampling, channel concat of the skip connection, residual
add; reference varnet.py:98,116,139-146,176-181, unet.py:6-24,119-140) of the *producing* layers is
applied while the operand of the *consuming* convolution is staged as 16-bit (hi, lo) pair tiles, so normalised /
activated / concatenated / pooled / summed tensors are never written to HBM.  What crosses an autograd
edge is always a raw fp32 conv output; its per-plane statistics ride along in a ``Raw`` handle and their
gradient is folded in analytically (the InstanceNorm / BatchNorm backward is linear in the incoming
gradient, so every consumer adds its own contribution).
"""
import ctypes
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import call, lib

MODE_DIRECT, MODE_POOL, MODE_D2S, MODE_UP = 0, 1, 2, 3
_IN_EPS = 1e-5

# 16-bit pair formats of the staged operands (include/san_b200.h).  Default: fp16 pairs everywhere = fp32-class
# products at the cost of the same 3 MMAs: forward operands (normalised activations, network inputs, weights) with
# static power-of-two scales, gradients (dY) with a dynamic one derived on the device from max|dY| (san_absmax).
# The operands of one MMA must share the format (a mixed f16 x bf16 tcgen05 MMA faults on the B200).
# SAN_TC_FMT selects the A/B experiments:
#   f16 (default) | f16nomix (fp16 pairs in the forward only; the backward re-stages X and stages dY / W as bf16
#   pairs: no dynamic scale, one more staging pass per conv) | bf16 (round-1 arithmetic: bf16 pairs everywhere).
FMT_BF16, FMT_F16 = 0, 1
_FMT_MODE = os.environ.get("SAN_TC_FMT", "f16")
assert _FMT_MODE in ("f16", "f16nomix", "bf16"), _FMT_MODE
_FMT_FWD = FMT_BF16 if _FMT_MODE == "bf16" else FMT_F16          # staged X and W of the forward conv
_FMT_BWD = FMT_F16 if _FMT_MODE == "f16" else FMT_BF16           # staged dY, X and W of the backward GEMMs
# InstanceNorm backward as one kernel per tensor (san_in_bwd_fused_map); SAN_IN_BWD_FUSED=0: the three-kernel path
_IN_BWD_FUSED = os.environ.get("SAN_IN_BWD_FUSED", "1") != "0"

# Weight-gradient GEMM on a side stream, overlapped with the HBM-bound element-wise backward of the same layer
# (normalisation backward of the conv's inputs + staging of their gradient for the producing layer): the GEMM is
# tensor / shared-memory bound and its persistent CTA (192 threads) leaves room on every SM for those kernels' CTAs.
# Two tcgen05 kernels never co-reside (each CTA allocates all 512 TMEM columns), so the fork happens AFTER the data
# gradient of the layer and the join BEFORE its backward returns (autograd accumulates dW on the main stream, and the
# next tcgen05 launch is the producing layer's data gradient).  SAN_WG_OVERLAP=0: everything on one stream.
_WG_OVERLAP = os.environ.get("SAN_WG_OVERLAP", "1") != "0"
_WG_OVERLAP_MIN_ELEMS = 1 << 23
_WG_PRESTAGE = os.environ.get("SAN_WG_PRESTAGE", "1") != "0"      # 0: the producing layer stages its dY itself (A/B runs)
# InstanceNorm statistics of a conv output from the conv's own epilogue (san_tc_conv_stats); SAN_EPI_STATS=0: the separate
# san_plane_stats_in pass over the tensor
_EPI_STATS = os.environ.get("SAN_EPI_STATS", "1") != "0"
_side = {}


def _side_stream(device):
    """The side stream that belongs to the CURRENT stream (one per sub-batch chain of varnet.VarNet._forward_streams)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=idx)
    return _side[key]


# Staged operands are as large as the activations they come from.  They are re-created in the backward
# pass (for the weight gradient) unless HBM is plentiful: while live tensors take less than
# this fraction of the device memory, the forward keeps them.  A B200 has 180 GB; the benchmark step (bs 64) needs 65 GB
# without them and 140 GiB with all of them kept, which is what 0.75 amounts to there (measured, profiles/r2v_*: 371.1 ms
# per step at 0.55 = 120 GiB peak, 365.5 at 0.65, 361.6 at 0.75 = 140 GiB).  SAN_KEEP_STAGED_BELOW overrides.
KEEP_STAGED_BELOW = float(os.environ.get("SAN_KEEP_STAGED_BELOW", "0.75"))
_total_mem = {}


# BatchNorm batch statistics under data parallelism: None = per-rank statistics (standard DDP semantics); a
# torch.distributed process group (set by parallel.attach(..., sync_bn=True)) = statistics of the GLOBAL batch: the
# per-plane sums of every rank are all-gathered (a few hundred floats per layer) before the finalise kernels, so a
# sharded step equals the single-process reference at the global batch size (reference unet.py:119-140 runs one process).
SYNC_BN_GROUP = None


def _gather_planes(t):
    """[planes] per-(n, c) statistics of this rank -> [world * planes] of the global batch (rank-major = sample-major)."""
    import torch.distributed as dist
    w = dist.get_world_size(SYNC_BN_GROUP)
    out = torch.empty(w * t.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=SYNC_BN_GROUP)
    return out, w


def _keep_staged(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _total_mem:
        _total_mem[idx] = torch.cuda.get_device_properties(idx).total_memory
    return torch.cuda.memory_allocated(idx) < KEEP_STAGED_BELOW * _total_mem[idx]


class _StageTerm(ctypes.Structure):
    """``san_stage_term`` of include/san_b200.h."""
    _fields_ = [("y", ctypes.c_void_p), ("mu", ctypes.c_void_p), ("a", ctypes.c_void_p), ("b", ctypes.c_void_p),
                ("slope", ctypes.c_float), ("C", ctypes.c_int), ("mode", ctypes.c_int), ("accumulate", ctypes.c_int)]


class Raw:
    """One term of a conv operand: a raw fp32 NCHW tensor + how a consumer reads it.

    ``norm``: None | 'in' (InstanceNorm2d, biased variance, eps 1e-5, varnet.py:141) | 'bn'
    (``bn`` = the ``nn.BatchNorm2d`` whose affine parameters / running buffers apply, unet.py:125);
    ``slope``: LeakyReLU slope (1 = none); ``d2s``: the tensor is the [N, 4C, h, w] output of the 1x1 form
    of ConvTranspose2d(2, stride 2), normalised over all four sub-planes of a channel; ``up``: the tensor
    is read through nearest x2 up-sampling (BatchNorm statistics of the up-sampled map = those of the
    low-resolution map with 4x the element count)."""

    def __init__(self, y, norm=None, slope=1.0, d2s=False, bn=None, up=False):
        assert y.dtype == torch.float32 and y.dim() == 4
        assert (norm == "bn") == (bn is not None)
        self.y = y if y.is_contiguous() else y.contiguous()
        self.norm, self.slope, self.d2s, self.bn, self.up = norm, float(slope), bool(d2s), bn, bool(up)
        self._coef = None

    @property
    def channels(self):
        return self.y.shape[1] // 4 if self.d2s else self.y.shape[1]

    def planes(self):
        N, C, H, W = self.y.shape
        return (N * C // 4, 4 * H * W) if self.d2s else (N * C, H * W)

    def coef(self):
        """Per-plane coefficient table, computed once per raw tensor and shared by all its consumers:
        'in': [4, planes] = mean, m2, a (= rstd), b (= 0);  'bn': [6, planes] = mean, m2, mu, a, b, sa."""
        if self.norm is None:
            return None
        if self._coef is None:
            planes, P = self.planes()
            yd = self.y.detach()
            if self.norm == "in":
                st = torch.empty(4, planes, dtype=torch.float32, device=yd.device)
                sums = getattr(self.y, "_san_sums", None)
                if sums is not None:
                    # the producing conv's epilogue already accumulated the per-plane sums (san_tc_conv_stats)
                    group = 4 if self.d2s else 1
                    call("in_stats_from_sums", sums, st[0], st[1], st[2], st[3], planes, group, P // group, _IN_EPS)
                else:
                    call("plane_stats_in", yd, st[0], st[1], st[2], st[3], planes, P, _IN_EPS)
            else:
                bn = self.bn
                N, C = yd.shape[0], yd.shape[1]
                training = bn.training or not bn.track_running_stats
                st = torch.empty(6, planes, dtype=torch.float32, device=yd.device)
                if training:
                    call("plane_stats", yd, st[0], st[1], planes, P)
                    if bn.track_running_stats:
                        bn.num_batches_tracked.add_(1)
                    if self.up:
                        st[1].mul_(4.0)      # m2 of the up-sampled map
                if training and SYNC_BN_GROUP is not None:
                    g0, w = _gather_planes(st[0])
                    g1, _ = _gather_planes(st[1])
                    out = torch.empty(4, w * planes, dtype=torch.float32, device=yd.device)
                    call("bn_finalize_fwd", g0, g1, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                         out[0], out[1], out[2], out[3], N * w, C, P * (4 if self.up else 1), bn.eps, bn.momentum, 1)
                    st[2:6].copy_(out[:, :planes])          # the coefficients are per channel: identical for every sample
                else:
                    call("bn_finalize_fwd", st[0], st[1], bn.weight, bn.bias, bn.running_mean, bn.running_var,
                         st[2], st[3], st[4], st[5], N, C, P * (4 if self.up else 1), bn.eps, bn.momentum, int(training))
            self._coef = st
        return self._coef


def _staged_act(N, H, W, C, device):
    return torch.empty(lib().san_tc_staged_act_elems(N, H, W, C), dtype=torch.bfloat16, device=device)


def _stage(xs, N, H, W, Cpad, terms, fmt, absmax=None):
    """terms: list of (y, mu, a, b, slope, C, mode, accumulate); fmt: FMT_BF16 | FMT_F16; absmax: device scalar
    max|y| selecting the dynamic fp16-pair scale (gradient operands)."""
    arr = (_StageTerm * len(terms))()
    for t, (y, mu, a, b, slope, C, mode, acc) in zip(arr, terms):
        t.y = y.data_ptr()
        t.mu = mu.data_ptr() if mu is not None else None
        t.a = a.data_ptr() if a is not None else None
        t.b = b.data_ptr() if b is not None else None
        t.slope, t.C, t.mode, t.accumulate = slope, C, mode, int(acc)
    call("tc_stage_terms", xs, N, H, W, Cpad, ctypes.addressof(arr), len(terms), fmt, absmax)


def _stage_weights_now(w, ws, dgrad, H, W, fmt):
    Cout, Cin, K, _ = w.shape
    call("tc_stage_weights", w, ws, H, W, Cout, Cin, K, int(dgrad), fmt)


def _alloc_staged_weights(w, dgrad, H, W):
    Cout, Cin, K, _ = w.shape
    n = lib().san_tc_staged_weight_elems(H, W, Cin if dgrad else Cout, Cout if dgrad else Cin, K)
    assert n > 0, f"tcgen05 conv: unsupported shape H={H} W={W} {tuple(w.shape)}"
    return torch.empty(n, dtype=torch.bfloat16, device=w.device)


# Staged weights of nn.Parameters are CACHED per (parameter, forward / data-gradient form, image size, pair format) and
# keyed on the parameter's version counter (bumped by every in-place update: torch's optimisers, load_state_dict, and
# optim.AdamW / parallel.attach of this package call torch.autograd.graph.increment_version).  A training step changes
# every weight once, so it re-stages every weight once either way - but not as 650 tiny launches in the dependency chain
# of the convolutions: the first stale weight a step meets re-stages ALL stale entries on the side stream, in the order
# the step will use them, while the main stream goes on; a consumer waits on the event of its batch of 24.  Evaluation
# (weights fixed) stages nothing at all.  Non-parameter weights (the reshaped ConvTranspose2d filter, spectrally
# normalised GAN weights) and CUDA-graph captures keep the inline launch.  SAN_WS_CACHE=0 disables the cache.
_WS_CACHE = os.environ.get("SAN_WS_CACHE", "1") != "0"
_ws_entries = {}


class _WsEntry:
    __slots__ = ("wref", "shape", "ptr", "dgrad", "H", "W", "fmt", "ws", "version", "event")


def invalidate_weight_cache():
    """Forget every cached staged weight (after mutating parameters behind autograd's back, e.g. through ``.data``)."""
    _ws_entries.clear()


def _restage_stale(device):
    main, side = torch.cuda.current_stream(), _side_stream(device)
    side.wait_stream(main)              # after every kernel that still reads the old staged weights / writes the new values
    dead, batch = [], []
    with torch.cuda.stream(side):
        for key, e in _ws_entries.items():
            w = e.wref()
            if w is None:
                dead.append(key)
                continue
            if e.version == w._version or w.device != device or e.ptr != w.data_ptr():
                continue
            _stage_weights_now(w, e.ws, e.dgrad, e.H, e.W, e.fmt)
            e.version = w._version
            batch.append(e)
            if len(batch) == 24:
                ev = torch.cuda.Event()
                ev.record(side)
                for b in batch:
                    b.event = ev
                batch = []
        if batch:
            ev = torch.cuda.Event()
            ev.record(side)
            for b in batch:
                b.event = ev
    for key in dead:
        del _ws_entries[key]


def _stage_weights(w, dgrad, H, W, fmt):
    import weakref
    if not (_WS_CACHE and isinstance(w, torch.nn.Parameter)) or torch.cuda.is_current_stream_capturing():
        ws = _alloc_staged_weights(w, dgrad, H, W)
        _stage_weights_now(w, ws, dgrad, H, W, fmt)
        return ws
    key = (id(w), bool(dgrad), H, W, fmt)
    e = _ws_entries.get(key)
    if e is None or e.wref() is not w or e.shape != tuple(w.shape) or e.ptr != w.data_ptr():
        # (a different storage behind the same Parameter object - module.to(), p.data = ... - is a new entry)
        e = _WsEntry()
        e.wref, e.shape, e.ptr, e.dgrad, e.H, e.W, e.fmt = weakref.ref(w), tuple(w.shape), w.data_ptr(), bool(dgrad), H, W, fmt
        e.ws = _alloc_staged_weights(w, dgrad, H, W)
        _stage_weights_now(w, e.ws, dgrad, H, W, fmt)
        e.version = w._version
        e.event = torch.cuda.Event()        # other streams (sub-batch chains) wait for this first staging too
        e.event.record()
        _ws_entries[key] = e
        return e.ws
    if e.version != w._version:
        _restage_stale(w.device)
    if e.event is not None:             # (kept: every stream that uses the entry waits for the sweep that staged it)
        torch.cuda.current_stream().wait_event(e.event)
    return e.ws


def _pad8(c):
    """Channels of a staged tensor: ceil(C / 8) groups of 8 (no all-zero groups)."""
    return (c + 7) // 8 * 8


def _coef_views(norm, st):
    """-> (mu, a, b, sa) rows of a coefficient table."""
    if st is None:
        return None, None, None, None
    if norm == "in":
        return st[0], st[2], None, st[2]
    return st[2], st[3], st[4], st[5]


def _tag_absmax(dy):
    """Device scalar that the kernel writing ``dy`` fills with max|dy|; it rides on the tensor object to the producing
    layer's backward (autograd hands the very same tensor over when the gradient has a single consumer).  The tensor
    version is recorded: an in-place accumulation by the autograd engine invalidates the tag."""
    if _FMT_BWD != FMT_F16:
        return None
    amax = torch.empty(1, dtype=torch.float32, device=dy.device)
    dy._san_absmax = (amax, dy._version)
    return amax


def _known_absmax(gy):
    tag = getattr(gy, "_san_absmax", None)
    if tag is not None and tag[1] == gy._version and gy.is_contiguous():
        return tag[0]
    return None


def _prestage(dy, t):
    """``dy`` = the gradient this layer's backward just wrote for a conv output with a single consumer: it IS the ``gy`` of
    the producing conv's backward, so it is staged here - inside the window in which this layer's weight-gradient GEMM runs
    on the side stream - and rides on the tensor object like the absmax tag."""
    if not (_WG_OVERLAP and _WG_PRESTAGE and t["prestage"] and _FMT_BWD == FMT_F16):
        return
    tag = getattr(dy, "_san_absmax", None)
    if tag is None:
        return
    N, C, H, W = dy.shape
    gys = _staged_act(N, H, W, C, dy.device)
    _stage(gys, N, H, W, _pad8(C), [(dy, None, None, None, 1.0, C, MODE_DIRECT, False)], _FMT_BWD, tag[0])
    dy._san_staged = (gys, tag[0], dy._version)


def _known_staged(gy):
    tag = getattr(gy, "_san_staged", None)
    if tag is not None and tag[2] == gy._version and gy.is_contiguous():
        return tag[0], tag[1]
    return None, None


class _FusedConv(Function):
    """y = conv2d(concat_k sum_j act(norm(resample(raw_kj))), w) + bias on the tcgen05 kernels.

    inputs: w, bias, spec, then per term: y [, gamma, beta when the term is BatchNorm-normalised]."""

    @staticmethod
    def forward(ctx, w, bias, spec, *tensors):
        K, metas, stats = spec  # metas[i] = (norm, slope, d2s, mode, accumulate, coef, bn_training, producer info);
                                # stats: None | [] = holder that receives the epilogue's per-plane sums of the output
        w = w.contiguous()
        Cout, Cin, Kw, _ = w.shape
        assert Kw == K
        ys, ti = [], 0
        for m in metas:
            ys.append(tensors[ti])
            ti += 3 if m[0] == "bn" else 1
        y0, m0 = ys[0], metas[0]
        N = y0.shape[0]
        if m0[3] == MODE_POOL:
            H, W = y0.shape[2] // 2, y0.shape[3] // 2
        elif m0[3] in (MODE_D2S, MODE_UP):
            H, W = y0.shape[2] * 2, y0.shape[3] * 2
        else:
            H, W = y0.shape[2], y0.shape[3]
        terms, ctot = [], 0
        for y, (norm, slope, d2s, mode, acc, coef, _, _) in zip(ys, metas):
            C = y.shape[1] // 4 if d2s else y.shape[1]
            mu, a, b, _sa = _coef_views(norm, coef)
            terms.append((y, mu, a, b, slope, C, mode, acc))
            if not acc:
                ctot += C
        assert ctot == Cin, (ctot, Cin)
        xs = _staged_act(N, H, W, Cin, w.device)
        _stage(xs, N, H, W, _pad8(Cin), terms, _FMT_FWD)
        ws = _stage_weights(w, False, H, W, _FMT_FWD)
        out = torch.empty(N, Cout, H, W, dtype=torch.float32, device=w.device)
        if stats is not None and _EPI_STATS and lib().san_tc_conv_stats_supported(H, W, Cin, Cout, K):
            sums = torch.empty(2 * N * Cout, dtype=torch.float64, device=w.device)
            call("tc_conv_stats", xs, ws, bias, out, N, H, W, Cin, Cout, K, 0, 3 * _FMT_FWD, None, sums)
            stats.append(sums)
        else:
            call("tc_conv", xs, ws, bias, out, N, H, W, Cin, Cout, K, 0, 3 * _FMT_FWD, None)
        # The staged operand is kept for the backward only while HBM is plentiful (_keep_staged); otherwise
        # the weight gradient re-stages it from the raw tensors, which autograd holds anyway for the
        # normalisation backward.
        coefs = [m[5] for m in metas if m[5] is not None]
        keep = (_keep_staged(w.device) and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
                and _FMT_BWD == _FMT_FWD)    # f16nomix: the weight gradient wants bf16 pairs -> re-staged in backward
        ctx.save_for_backward(w, *tensors, *coefs, *([xs] if keep else []))
        ctx.meta = (K, [(m[0], m[1], m[2], m[3], m[4], m[5] is not None, m[6], m[7]) for m in metas], (N, H, W),
                    bias is not None, len(tensors), keep)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        K, metas, (N, H, W), has_bias, ntens, kept = ctx.meta    # (the forward's `stats` holder is not needed here)
        saved = ctx.saved_tensors
        w = saved[0]
        tensors = saved[1:1 + ntens]
        coef_list = list(saved[1 + ntens:len(saved) - (1 if kept else 0)])
        xs_kept = saved[-1] if kept else None
        Cout, Cin = w.shape[0], w.shape[1]
        gy_in = gy
        gy = gy if gy.is_contiguous() else gy.contiguous()
        dev = w.device
        # unpack terms
        terms, ti, ci, c_next, c_prev = [], 0, 0, 0, 0
        for (norm, slope, d2s, mode, acc, has_coef, bn_training, prod) in metas:
            y = tensors[ti]
            gamma = tensors[ti + 1] if norm == "bn" else None
            C = y.shape[1] // 4 if d2s else y.shape[1]
            st = coef_list[ci] if has_coef else None
            ci += int(has_coef)
            c0 = c_prev if acc else c_next
            c_prev = c0
            if not acc:
                c_next = c0 + C
            terms.append(dict(y=y, gamma=gamma, norm=norm, slope=slope, d2s=d2s, mode=mode, acc=acc, st=st, C=C, c0=c0,
                              ti=ti, bn_training=bn_training, prestage=prod[0] and prod[1][0] == 1))
            ti += 3 if norm == "bn" else 1
        # dY staged once as a (hi, lo) pair: the operand of both the data- and the weight-gradient GEMMs (already done by
        # the consuming layer's backward when this conv output had a single consumer: _prestage)
        gys, amax = _known_staged(gy_in)
        if gys is None:
            gys = _staged_act(N, H, W, Cout, dev)
            amax = None
            if _FMT_BWD == FMT_F16:          # dynamic power-of-two scale of the gradient operand
                amax = _known_absmax(gy_in)  # left by the consumer's normalisation backward when it wrote gy
                if amax is None:
                    amax = torch.empty(1, dtype=torch.float32, device=dev)
                    call("absmax", gy, gy.numel(), amax)
            _stage(gys, N, H, W, _pad8(Cout), [(gy, None, None, None, 1.0, Cout, MODE_DIRECT, False)], _FMT_BWD, amax)
        need_dgrad = any(ctx.needs_input_grad[3 + t["ti"]] for t in terms)
        need_wgrad = ctx.needs_input_grad[0] or (has_bias and ctx.needs_input_grad[1])
        # (small problems - the reference's batch of 4, the deepest layers - gain nothing from the fork and pay its host cost)
        overlap = _WG_OVERLAP and need_dgrad and need_wgrad and N * H * W * max(Cin, Cout) >= _WG_OVERLAP_MIN_ELEMS

        # ---- weight gradient: tcgen05 GEMM over the pixel dimension on the staged dY and the (kept or re-staged) input
        dw = db = None
        wg_done = None

        use_tc_wgrad = bool(lib().san_tc_wgrad_supported(H, W, Cin, Cout, K))
        x32 = None

        def weight_gradient():      # launches only: every buffer is allocated on the main stream beforehand
            if use_tc_wgrad:
                call("tc_wgrad", gys, xs, dw, db, gy if has_bias else None, N, H, W, Cin, Cout, K, 3 * _FMT_BWD, amax)
            else:   # tiny images (W < 16): fp32 CUDA-core kernel on the un-staged operand
                call("tc_unstage_act", xs, x32, N, Cin, H, W, _FMT_BWD)
                call("conv2d_wgrad", x32, gy, dw, db, N, Cin, H, W, Cout, K, 0, 0)

        xs = None
        if need_wgrad:
            if xs_kept is not None:
                xs = xs_kept
            else:
                xs = _staged_act(N, H, W, Cin, dev)
                _stage(xs, N, H, W, _pad8(Cin),
                       [(t["y"], *_coef_views(t["norm"], t["st"])[:3], t["slope"], t["C"], t["mode"], t["acc"]) for t in terms],
                       _FMT_BWD)
            dw = torch.empty_like(w)
            db = torch.empty(Cout, dtype=torch.float32, device=dev) if has_bias else None
            if not use_tc_wgrad:
                x32 = torch.empty(N, Cin, H, W, dtype=torch.float32, device=dev)
            if not overlap:
                weight_gradient()
        # ---- data gradient: the same tcgen05 conv on the staged dY with the flipped filter
        grads = [None] * ntens
        if need_dgrad:
            wsd = _stage_weights(w, True, H, W, _FMT_BWD)
            dx = torch.empty(N, Cin, H, W, dtype=torch.float32, device=dev)
            call("tc_conv", gys, wsd, None, dx, N, H, W, Cout, Cin, K, 0, 3 * _FMT_BWD, amax)
            if overlap:
                # fork: the weight gradient starts when the data gradient has finished and runs next to the element-wise
                # kernels below; every tensor it touches stays referenced until the join at the end of this function
                main, side = torch.cuda.current_stream(), _side_stream(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    weight_gradient()
                    wg_done = torch.cuda.Event()
                    wg_done.record(side)
            for t in terms:
                if not ctx.needs_input_grad[3 + t["ti"]]:
                    continue
                y, C, mode, slope, st, c0 = t["y"], t["C"], t["mode"], t["slope"], t["st"], t["c0"]
                Hy, Wy = y.shape[2], y.shape[3]
                if st is None and slope == 1.0:
                    # identity term: the gradient is the channel slice of dx, resampled back if need be
                    g = dx[:, c0:c0 + C]
                    if mode == MODE_DIRECT:
                        grads[t["ti"]] = g
                        continue
                    g = g if g.is_contiguous() else g.contiguous()
                    gs = torch.empty_like(y)
                    if mode == MODE_POOL:
                        call("up2", g, gs, N * C, H, W, 0.25)
                    elif mode == MODE_D2S:
                        call("space_to_depth2", g, gs, N, C, H // 2, W // 2)
                    else:
                        call("pool2", g, gs, N * C, H, W, 1.0)
                    grads[t["ti"]] = gs
                    continue
                # normalised / activated terms read their gradient IN PLACE from dx (channel offset + adjoint
                # of the resampling): no slice copy, no up2 / space_to_depth2 / pool2 temporaries
                dy = torch.empty_like(y)
                if st is None:      # bare LeakyReLU (cross.py:14)
                    planes = N * C
                    ones = torch.ones(planes, dtype=torch.float32, device=dev)
                    call("act_bwd_apply_map", dx, Cin, c0, mode, y, None, ones, None, slope, ones, None, None, dy,
                         N, C, Hy, Wy, _tag_absmax(dy))
                    _prestage(dy, t)
                    grads[t["ti"]] = dy
                    continue
                planes = st.shape[1]
                P = y.numel() // planes
                mu, a, b, sa = _coef_views(t["norm"], st)
                if t["norm"] == "in" and _IN_BWD_FUSED:
                    # per-plane statistics: reduce + coefficients + apply in one kernel, second pass out of L2
                    call("in_bwd_fused_map", dx, Cin, c0, mode, y, mu, a, slope, dy, N, C, Hy, Wy, _tag_absmax(dy))
                    _prestage(dy, t)
                    grads[t["ti"]] = dy
                    continue
                wk = torch.empty(5, planes, dtype=torch.float32, device=dev)
                call("act_bwd_reduce_map", dx, Cin, c0, mode, y, mu, a, b, sa, slope, wk[0], wk[1], N, C, Hy, Wy)
                if t["norm"] == "in":
                    call("in_finalize_bwd", wk[0], wk[1], a, wk[2], wk[3], wk[4], planes, P)
                else:
                    gamma = t["gamma"]
                    dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
                    if t["bn_training"] and SYNC_BN_GROUP is not None:
                        # global-batch statistics: the mean terms of dx use the sums of ALL ranks; dgamma / dbeta stay the
                        # local sums (the gradient all-reduce averages them like every other parameter gradient)
                        g0, w = _gather_planes(wk[0])
                        g1, _ = _gather_planes(wk[1])
                        out = torch.empty(3, w * planes, dtype=torch.float32, device=dev)
                        call("bn_finalize_bwd", g0, g1, gamma, sa, out[0], out[1], out[2], dgamma, dbeta,
                             y.shape[0] * w, y.shape[1], P, 1)
                        wk[2:5].copy_(out[:, :planes])
                        dbeta = wk[0].view(y.shape[0], -1).sum(0)
                        dgamma = wk[1].view(y.shape[0], -1).sum(0)
                    else:
                        call("bn_finalize_bwd", wk[0], wk[1], gamma, sa, wk[2], wk[3], wk[4], dgamma, dbeta,
                             y.shape[0], y.shape[1], P, int(t["bn_training"]))
                    grads[t["ti"] + 1], grads[t["ti"] + 2] = dgamma, dbeta
                call("act_bwd_apply_map", dx, Cin, c0, mode, y, mu, a, b, slope, wk[2], wk[3], wk[4], dy, N, C, Hy, Wy,
                     _tag_absmax(dy))
                _prestage(dy, t)
                grads[t["ti"]] = dy
        if wg_done is not None:
            torch.cuda.current_stream().wait_event(wg_done)     # join: dW is accumulated by autograd on the main stream
        return (dw, db, None, *grads)


def fused_conv(sources, weight, bias=None, modes=None, stats=False):
    """sources: list whose items are a ``Raw`` or a list of ``Raw`` (their SUM); the items are
    concatenated along channels.  modes: per-source resampling mode (default: direct; pixel shuffle for a
    ``d2s`` term; nearest x2 for an ``up`` term).  stats: the output will be InstanceNorm-normalised by its consumers:
    the conv epilogue accumulates its per-plane sums, so that no statistics pass re-reads the tensor (``Raw.coef``).
    Returns the raw fp32 conv output tensor."""
    metas, tensors = [], []
    for i, src in enumerate(sources):
        group = src if isinstance(src, (list, tuple)) else [src]
        for j, r in enumerate(group):
            mode = modes[i] if modes is not None else (MODE_D2S if r.d2s else MODE_UP if r.up else MODE_DIRECT)
            bn_training = bool(r.bn.training or not r.bn.track_running_stats) if r.norm == "bn" else False
            # consumer count of the raw tensor (shared, final by the time the backward runs) and whether a fused conv
            # produced it: a single-consumer conv output gets its gradient staged by THIS layer's backward
            uses = getattr(r.y, "_san_uses", None)
            if uses is None:
                uses = [0]
                r.y._san_uses = uses
            uses[0] += 1
            metas.append((r.norm, r.slope, r.d2s, mode, j > 0, r.coef(), bn_training,
                          (bool(getattr(r.y, "_san_conv_out", False)), uses)))
            tensors.append(r.y)
            if r.norm == "bn":
                tensors += [r.bn.weight, r.bn.bias]
    K = weight.shape[-1]
    holder = [] if stats else None
    out = _FusedConv.apply(weight, bias, (K, metas, holder), *tensors)
    out._san_conv_out = True
    if holder:
        out._san_sums = holder[0]
    return out
