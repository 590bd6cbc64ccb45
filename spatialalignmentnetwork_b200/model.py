"""``CSModel`` for the reconstruction + rec-guided registration path of the reference
(``reg`` in {'None', 'Rec'}; reference model.py:39-121, 142-169, 193-216, 265-321) on the
san_b200 kernels.  Same attribute protocol (``img_*`` / ``loss_*`` tensors harvested by
``get_vis``), same ``set_input`` / ``update`` / ``test`` / ``save`` / ``load`` calls, same
``net_T`` / ``net_R`` / ``net_mask`` checkpoint keys.  The GAN branches (``net_G`` / ``net_D``,
reg 'Mixed' / 'GAN-Only') are outside this hot path and raise.

Additions over the reference (none changes results): ``num_cascades`` config knob (reference
hard-codes 8, model.py:64), optional gradient all-reduce hook for one-process-per-GPU data
parallel training, optional per-cascade recomputation.
"""
import torch

from . import ops
from .basemodel import BaseModel, Config  # noqa: F401
from .cross import SpatialTransformer
from .masks import masks
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
        self.memo_init = (set(self.__dict__.keys()) | {"memo_init"}).copy()

    def build(self, cfg):
        super().build(cfg)
        assert cfg.lr == 1e-4
        self.net_mask = masks[cfg.mask](cfg.sparsity, cfg.shape)
        self.net_T = SpatialTransformer(channels=cfg.coils)
        self.net_R = VarNet(num_cascades=cfg.num_cascades if "num_cascades" in cfg else 8,
                            sens_chans=8, sens_pools=4, chans=18, pools=4, use_ref=True)
        if "checkpoint_cascades" in cfg:
            self.net_R.checkpoint_cascades = bool(cfg.checkpoint_cascades)
        self.optim_T = torch.optim.AdamW(self.net_T.parameters(), lr=cfg.lr, weight_decay=0)
        self.optim_R = torch.optim.AdamW(self.net_R.parameters(), lr=cfg.lr, weight_decay=0)
        self.use_amp = False  # fp32 parity path only

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

    def forwardT(self):
        moving = _cabs(self.img_aux)
        self.img_offset, self.img_grid = self.net_T(moving=moving, fixed=_cabs(self.img_sampled))
        self.img_warped = self.net_T.warp(moving, self.img_grid)
        self.img_warped_rss = rss(self.img_warped)
        self.loss_smooth = gradient_loss(self.img_offset)
        self.loss_all = self.loss_all + self.loss_smooth * self.cfg.weight_smooth

    def forwardR(self):
        self.img_rec = self.net_R(masked_kspace=self.img_k_sampled, mask=torch.logical_not(self.net_mask.pruned),
                                  ref=self.img_warped,
                                  num_low_frequencies=int(self.cfg.shape * self.cfg.sparsity * 0.32))
        self.loss_sim = ssimloss(self.img_full_rss, self.img_rec)
        self.loss_all = self.loss_all + self.loss_sim * self.cfg.weight_sim

    def _sync(self, nets):
        if self.grad_sync is not None:
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
            self.loss_all.backward()
            self._sync([self.net_R])
            self.optim_R.step()
        elif self.cfg.reg == "Rec":
            self.loss_all = 0
            self.forwardT()
            self.forwardR()
            self.optim_T.zero_grad()
            self.optim_R.zero_grad()
            self.loss_all.backward()
            self._sync([self.net_T, self.net_R])
            self.optim_T.step()
            self.optim_R.step()
        else:
            raise NotImplementedError(
                f"reg={self.cfg.reg!r}: the GAN branches (net_G / net_D) are outside the san_b200 hot path")

    def test(self):
        assert not self.training
        with torch.no_grad():
            self.loss_all = 0
            self.forwardT()
            self.loss_all = 0
            self.forwardR()
            mse = torch.mean((self.img_full_rss - self.img_rec) ** 2)
            self.metric_MSE = mse.item()
            self.metric_MAE = torch.mean((self.img_full_rss - self.img_rec).abs()).item()
            self.metric_PSNR = (10 * torch.log10(1.0 / mse)).item()       # metrics.py:35-38, range 1
            self.metric_SSIM = 1.0 - self.loss_sim.item()
        return -self.metric_PSNR

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
