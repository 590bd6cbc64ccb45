"""Parity of the whole ``reg='Rec'`` training step at the HEADLINE configurations (BASELINE.json cfg2 and cfg4), every
parameter gradient included: the CUDA path against the CPU oracle in fp32 (the reference's arithmetic) and fp64
(noise-floor calibration).  Reference: varnet.py:465-486, cross.py:23-38, model.py:142-169, 206-216.

Bars (north star: 1e-3 relative, fp32):
  * forward (img_rec, img_warped, img_offset, loss_all): < 1e-3 against the fp32 oracle;
  * all parameter gradients concatenated, against fp64: within 3x the distance of the CPU fp32 oracle from fp64
    (fp32 evaluation of 12 cascades of normalised LeakyReLU U-Nets has its own noise floor; the report prints both);
  * PSNR of img_rec against the reference's img_rec > 60 dB, SSIM > 0.9999 (BASELINE.json metric "PSNR/SSIM vs ref").
The full report is written to gpurun_out/ (copied to profiles/ when committed) and embedded by bench.py as `parity`."""
import json
import os
import random

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(tag, n, coils, H, W, cascades, seed=20221017):
    from oracle import parity
    from spatialalignmentnetwork_b200 import model as M
    torch.manual_seed(seed)
    random.seed(seed)
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=W, coils=coils, reg="Rec", mask="equispaced",
                   weight_smooth=1000.0, weight_sim=1.0, num_cascades=cascades,
                   gan_layers_G=[4, 8, 8], gan_layers_D=[[4, 4], [8, 8]])      # GAN nets unused by the Rec step: tiny
    net = M.CSModel(cfg)
    with torch.no_grad():
        torch.nn.init.normal_(net.net_T.net[-1].weight, 0, 1e-2)               # non-trivial warp
    g = torch.Generator().manual_seed(seed + 1)
    full = torch.complex(torch.rand(n, coils, H, W, generator=g), torch.rand(n, coils, H, W, generator=g))
    aux = torch.complex(torch.rand(n, coils, H, W, generator=g), torch.rand(n, coils, H, W, generator=g))
    sd_T = {k: v.clone() for k, v in net.net_T.state_dict().items()}
    sd_R = {k: v.clone() for k, v in net.net_R.state_dict().items()}
    pruned = net.net_mask.pruned.clone()
    net.to("cuda").train()
    net.set_input(full.cuda(), aux.cuda())
    net.loss_all = 0
    net.forwardT()
    net.forwardR()
    net.loss_all.backward()
    torch.cuda.synchronize()
    out = {k: getattr(net, k).detach().cpu() for k in ("img_rec", "img_warped", "img_offset")}
    out["loss_all"] = net.loss_all.item()
    grads = {"T." + k: p.grad.detach().cpu() for k, p in net.net_T.named_parameters() if p.grad is not None}
    grads.update({"R." + k: p.grad.detach().cpu() for k, p in net.net_R.named_parameters() if p.grad is not None})
    rep = parity.rec_step_report(sd_T, sd_R, full, aux, pruned, W, 0.25, cascades, out, grads)
    rep["config"]["tag"] = tag
    print(json.dumps(rep))
    od = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(od):
        with open(os.path.join(od, f"parity_{tag}.json"), "w") as f:
            json.dump(rep, f, indent=1)
    fw, gr, im = rep["forward"], rep["grad"]["all_concatenated"], rep["image_metrics"]
    assert not rep["grad"]["missing"], rep["grad"]["missing"]
    for k, v in fw.items():
        assert v["vs_fp32"] < 1e-3, (k, v)
    assert gr["vs_fp64"] < max(3.0 * gr["fp32_vs_fp64"], 2e-3), gr
    assert gr["cosine_vs_fp64"] > gr["fp32_cosine_vs_fp64"] - 0.01, gr          # same direction as well as the CPU fp32 oracle
    assert im["psnr_rec_vs_reference_rec_db"] > 60.0 and im["ssim_rec_vs_reference_rec"] > 0.9999, im
    return rep


def test_parity_cfg2_12_cascades_320():
    """BASELINE cfg2 geometry: 320x320, 12 cascades, single coil, 4x equispaced mask, 2 slices."""
    _run("cfg2_n2_320_c12", 2, 1, 320, 320, 12)


def test_parity_cfg4_15_coils_640x368():
    """BASELINE cfg4 geometry: 15 coils, 640x368, 4x equispaced mask over W = 368 (radix-23 FFT), 12 cascades, 1 slice."""
    _run("cfg4_n1_15c_640x368_c12", 1, 15, 640, 368, 12)
