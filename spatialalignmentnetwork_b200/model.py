"""``CSModel`` of the reference (model.py:39-321) on the san_b200 kernels: all four training modes
(``reg`` in {'None', 'Rec', 'Mixed', 'GAN-Only'}), ``test()`` with device-side metrics, same attribute
protocol (``img_*`` / ``loss_*`` / ``metric_*`` harvested by ``get_vis``), same ``set_input`` / ``update`` /
``test`` / ``save`` / ``load`` calls, same ``net_mask`` / ``net_G`` / ``net_D`` / ``net_T`` / ``net_R``
checkpoint keys and construction order (so a seeded build draws the same initial weights).

Additions over the reference (none changes results): ``num_cascades`` config knob (reference
hard-codes 8, model.py:64) and ``gan_layers_G`` / ``gan_layers_D`` (reference hard-codes the widths,
model.py:58-61), optional gradient all-reduce hook for one-process-per-GPU data parallel training,
optional per-cascade recomputation, ``fused_adamw`` (one multi-tensor AdamW launch per network instead of
torch's foreach kernels; same arithmetic).
"""
import math

import torch

from . import ops
from .basemodel import BaseModel, Config  # noqa: F401
from .cross import SpatialTransformer
from .gan import NetD, NetG, loss_gan
from .lnccloss import lncc_loss
from .masks import masks
from .miloss import ms_mi_loss
from .optim import AdamW
from .signal_utils import fft2, fftshift2, ifft2, rss
from .ssimloss import ssimloss
from .varnet import VarNet


def gradient_loss(s):
    """Mean squared finite differences of the displacement field (reference model.py:21-28)."""
    assert s.shape[-1] == 2, "not 2D grid?"
    return ops.GradientLoss.apply(s)


def _cabs(x):
    """|x| per element of a complex64 [N,C,H,W] tensor (``.abs()`` in model.py:144-149)."""
    if not torch.is_complex(x):
        return x.abs()
    N, C, H, W = x.shape
    return ops.Rss.apply(x.reshape(N * C, 1, H, W)).reshape(N, C, H, W)


class CSModel(BaseModel):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.grad_sync = None  # callable(list_of_params) installed by parallel.attach()
        self.grad_buckets = None  # parallel.GradBuckets (overlapped per-cascade all-reduce) installed by parallel.attach()
        self.memo_init = (set(self.__dict__.keys()) | {"memo_init"}).copy()

    def build(self, cfg):
        super().build(cfg)
        assert cfg.lr == 1e-4
        if cfg.mask in ("mask", "taylor"):                 # reference model.py:53-56 ('taylor' raises: out of scope)
            self.net_mask = masks[cfg.mask](cfg.shape)
        else:
            self.net_mask = masks[cfg.mask](cfg.sparsity, cfg.shape)
        # same construction order as the reference (model.py:58-71): identical RNG consumption under a seed
        layers_G = tuple(cfg.gan_layers_G) if "gan_layers_G" in cfg else (64, 128, 256, 512, 512)
        layers_D = ([list(b) for b in cfg.gan_layers_D] if "gan_layers_D" in cfg else
                    ([64] * 2, [128] * 2, [256] * 2, [256] * 2, [256] * 2))
        self.net_G = NetG(in_channels=1, out_channels=1, layers=layers_G)
        self.net_D = NetD(in_channels=2, layers=layers_D)
        self.net_T = SpatialTransformer(channels=cfg.coils)
        self.net_R = VarNet(num_cascades=cfg.num_cascades if "num_cascades" in cfg else 8,
                            sens_chans=8, sens_pools=4, chans=18, pools=4, use_ref=True)
        if "checkpoint_cascades" in cfg:
            self.net_R.checkpoint_cascades = bool(cfg.checkpoint_cascades)
        opt = AdamW if ("fused_adamw" not in cfg or cfg.fused_adamw) else torch.optim.AdamW
        self.optim_G = opt(self.net_G.parameters(), lr=cfg.lr, weight_decay=0)
        self.optim_D = opt(self.net_D.parameters(), lr=cfg.lr, weight_decay=0)
        self.optim_T = opt(self.net_T.parameters(), lr=cfg.lr, weight_decay=0)
        self.optim_R = opt(self.net_R.parameters(), lr=cfg.lr, weight_decay=0)
        self.optim_M = opt(self.net_mask.parameters(), lr=cfg.lr, weight_decay=0)
        self.use_amp = False  # fp32 parity path only (the reference's GradScaler is a no-op without amp)

    def set_input(self, img_full, img_aux=None):
        for name in [k for k in self.__dict__ if k.startswith(("loss_", "img_", "metric_"))]:
            delattr(self, name)
        self.img_full = img_full
        self.img_aux = torch.zeros_like(img_full) if img_aux is None else img_aux
        with torch.no_grad():
            keep = (1 - self.net_mask.pruned.float())
        self.img_k_full = fft2(self.img_full)
        with torch.no_grad():
            self.img_k_sampled = self.img_k_full * keep            # multiply, model.py:113
        self.img_sampled = ifft2(self.img_k_sampled)
        self.img_full_rss = rss(self.img_full)
        self.img_sampled_rss = rss(self.img_sampled)
        self.img_aux_rss = rss(self.img_aux)
        with torch.no_grad():
            self.img_mask = fftshift2(torch.ones_like(self.img_full_rss) - self.net_mask.pruned.float())

    def forwardG(self):
        """Modality translation (reference model.py:123-140): the first half of the batch is warped then
        translated (R -> TR), the second half translated then warped (T -> RT)."""
        aux_TR, aux_RT = torch.chunk(self.img_aux_rss, 2, dim=0)
        T = self.net_G(aux_RT)
        R, RT = torch.chunk(self.net_T.warp(img=torch.cat((aux_TR, T)), grid=self.img_grid), 2)
        TR = self.net_G(R)
        self.img_synth = torch.cat((R, T), dim=0)
        self.img_aligned = torch.cat((TR, RT), dim=0)
        self.loss_gan_sim = ops.l1_loss(self.img_aligned, self.img_full_rss)
        self.loss_all = self.loss_all + self.loss_gan_sim * self.cfg.weight_gan_sim

    def forwardD(self, D_loss):
        """Hinge GAN terms on [image, 0] pairs (reference model.py:171-190); the zero second channel is read as
        a second conv source instead of being concatenated."""
        zeros = torch.zeros_like(self.img_full_rss)
        if D_loss:
            self.loss_gan_Dfake = loss_gan(self.net_D.forward_sources([self.img_aligned.detach(), zeros]),
                                           real=False, D_loss=True)
            self.loss_gan_Dreal = loss_gan(self.net_D.forward_sources([self.img_full_rss.detach(), zeros]),
                                           real=True, D_loss=True)
            self.loss_all = self.loss_all + (self.loss_gan_Dfake + self.loss_gan_Dreal) * self.cfg.weight_gan
        else:
            self.loss_gan_G = loss_gan(self.net_D.forward_sources([self.img_aligned, zeros]), real=False, D_loss=False)
            self.loss_all = self.loss_all + self.loss_gan_G * self.cfg.weight_gan

    def forwardT(self):
        moving = _cabs(self.img_aux)
        self.img_offset, self.img_grid = self.net_T(moving=moving, fixed=_cabs(self.img_sampled))
        self.img_warped = self.net_T.warp(moving, self.img_grid)
        self.img_warped_rss = rss(self.img_warped)
        self.loss_smooth = gradient_loss(self.img_offset)
        self.loss_all = self.loss_all + self.loss_smooth * self.cfg.weight_smooth
        # Optional registration similarity terms of BASELINE configs 3 and 5 (the reference carries both losses,
        # lnccloss.py / miloss.py, but its live path has them commented out, model.py:12): absent unless the
        # config names a weight.
        if "weight_lncc" in self.cfg and self.cfg.weight_lncc:
            self.loss_lncc = lncc_loss(self.img_full_rss, self.img_warped_rss)
            self.loss_all = self.loss_all + self.loss_lncc * self.cfg.weight_lncc
        if "weight_mi" in self.cfg and self.cfg.weight_mi:
            self.loss_mi = ms_mi_loss(self.img_full_rss, self.img_warped_rss)
            self.loss_all = self.loss_all + self.loss_mi * self.cfg.weight_mi

    def forwardR(self):
        self.img_rec = self.net_R(masked_kspace=self.img_k_sampled, mask=torch.logical_not(self.net_mask.pruned),
                                  ref=self.img_warped,
                                  num_low_frequencies=int(self.cfg.shape * self.cfg.sparsity * 0.32))
        self.loss_sim = ssimloss(self.img_full_rss, self.img_rec)
        self.loss_all = self.loss_all + self.loss_sim * self.cfg.weight_sim

    def _names(self, nets):
        return [k for n in nets for k, v in self.__dict__.items() if v is n]

    def _arm(self, nets):
        """Before ``backward()``: route the gradients of ``nets`` into the flat all-reduce buckets (parallel.GradBuckets),
        whose exchange then overlaps the rest of the backward.  No-op on one GPU."""
        if self.grad_buckets is not None:
            self.grad_buckets.arm(self._names(nets))

    def _sync(self, nets):
        """After ``backward()``: gradients of ``nets`` are the mean over ranks when this returns."""
        if self.grad_buckets is not None:
            self.grad_buckets.sync(self._names(nets))
        elif self.grad_sync is not None:
            self.grad_sync([p for n in nets for p in n.parameters()])

    def update(self):
        assert self.training
        if self.cfg.reg == "None":
            self.loss_all = 0
            with torch.no_grad():
                self.forwardT()
            self.loss_all = 0
            self.forwardR()
            self.optim_R.zero_grad()
            self._arm([self.net_R])
            self.loss_all.backward()
            self._sync([self.net_R])
            self.optim_R.step()
        elif self.cfg.reg == "Rec":
            self.loss_all = 0
            self.forwardT()
            self.forwardR()
            self.optim_T.zero_grad()
            self.optim_R.zero_grad()
            self._arm([self.net_T, self.net_R])
            self.loss_all.backward()
            self._sync([self.net_T, self.net_R])
            self.optim_T.step()
            self.optim_R.step()
        elif self.cfg.reg in ("Mixed", "GAN-Only"):
            # (reconstruction and) GAN-guided registration: update T, G (and R), then D (model.py:217-260)
            with_R = self.cfg.reg == "Mixed"
            self.loss_all = 0
            self.forwardT()
            self.forwardG()
            if with_R:
                self.forwardR()
            self.forwardD(D_loss=False)
            nets = [self.net_T, self.net_G] + ([self.net_R] if with_R else [])
            opts = [self.optim_T, self.optim_G] + ([self.optim_R] if with_R else [])
            for o in opts:
                o.zero_grad()
            self._arm(nets)
            self.loss_all.backward()
            self._sync(nets)
            for o in opts:
                o.step()
            self.loss_all = 0
            self.forwardD(D_loss=True)
            self.optim_D.zero_grad()
            self._arm([self.net_D])
            self.loss_all.backward()
            self._sync([self.net_D])
            self.optim_D.step()
        else:
            assert False, f"unknown reg {self.cfg.reg!r}"
        del self.loss_all          # reference model.py:261: the graph is released, get_vis() does not log it

    def test(self):
        assert not self.training
        with torch.no_grad():
            self.loss_all = 0
            self.forwardT()
            self.loss_all = 0
            self.forwardG()
            self.loss_all = 0
            self.forwardR()
            # metrics.py on the device: fp64 reductions, a few scalars cross to the host
            n = self.img_full_rss.numel()
            se, ae, _ = ops.error_sums(self.img_full_rss, self.img_rec)
            self.metric_MI = ops.mi_metric(self.img_full_rss, self.img_warped_rss)
            self.metric_MSE = se / n
            self.metric_MAE = ae / n
            self.metric_PSNR = 10.0 * math.log10(1.0 / self.metric_MSE)     # metrics.py:35-38, data_range 1
            self.metric_SSIM = 1.0 - self.loss_sim.item()                   # skimage SSIM == 1 - ssimloss
        return -self.metric_MI if self.cfg.reg == "GAN-Only" else -self.metric_PSNR

    def get_vis(self, content=None):
        assert content in [None, "scalars", "histograms", "images"]
        vis = {}
        if content in ("scalars", None):
            vis["scalars"] = {}
            for k, v in self.__dict__.items():
                if k.startswith("loss_") and v is not None:
                    vis["scalars"][k] = v.detach().item() if torch.is_tensor(v) else v
                if k.startswith("metric_") and v is not None:
                    vis["scalars"][k] = v
        if content in ("images", None):
            vis["images"] = {k: v.detach() for k, v in self.__dict__.items()
                             if k.startswith("img_") and v is not None and v.shape[1] in (1, 3)
                             and not torch.is_complex(v)}
        if content in ("histograms", None):
            vis["histograms"] = {"weights": {"values": self.net_mask.weight.detach()}}
        return vis
