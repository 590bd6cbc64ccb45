"""The drop-in recipe of INTEGRATION.md (A), exercised with the UNMODIFIED reference ``model.py`` in the driver's
seat: this repository's modules are aliased into ``sys.modules`` under the reference's flat module names, then the
reference's own ``CSModel`` builds its networks from them and runs ``set_input`` / ``forwardT`` / ``forwardR`` /
``forwardG`` / ``forwardD`` / ``update()``.  On CPU the C-ABI ops are the torch stand-ins of tests/emulation.py, so
this checks the module API (names, constructor arguments, call signatures, attribute protocol), not the kernels.
Needs /root/reference (build container); skipped elsewhere."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("SAN_REFERENCE", "/root/reference")

_SCRIPT = r'''
import importlib, random, sys, warnings
import torch
warnings.filterwarnings("ignore")
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, root + "/tests", root + "/tests/golden"]
import emulation, model_shim                      # op stand-ins; skimage stub for the reference's metrics.py
class MP:
    def setattr(self, o, n, v): setattr(o, n, v)
emulation.install(MP())
from spatialalignmentnetwork_b200 import unet as U, varnet as V
V.USE_TC = U.USE_TC = True
for name in ("signal_utils", "varnet", "cross", "unet", "gan", "augment", "ssimloss", "lnccloss", "miloss"):   # INTEGRATION.md (A)
    sys.modules[name] = importlib.import_module("spatialalignmentnetwork_b200." + name)
sys.path.insert(0, ref)
import model as refmodel                          # the reference's own model.py
from basemodel import Config                      # ... and basemodel.py
from conftest import load_golden, rel_l2, sub
assert refmodel.__file__.startswith(ref)
assert refmodel.VarNet.__module__.startswith("spatialalignmentnetwork_b200")
assert refmodel.NetG.__module__.startswith("spatialalignmentnetwork_b200")

def build(reg, seed, g_layers, d_layers):
    cfg = Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg=reg, mask="equispaced", weight_smooth=1000.0,
                 weight_gan=0.1, weight_gan_sim=1.0, weight_sim=1.0, use_amp=False)
    random.seed(seed)
    oV, oG, oD = refmodel.VarNet, refmodel.NetG, refmodel.NetD
    refmodel.VarNet = lambda **kw: oV(**{**kw, "num_cascades": 2, "chans": 4, "pools": 2, "sens_chans": 2, "sens_pools": 2})
    refmodel.NetG = lambda **kw: oG(**{**kw, "layers": g_layers})
    refmodel.NetD = lambda **kw: oD(**{**kw, "layers": d_layers})
    net = refmodel.CSModel(cfg)
    refmodel.VarNet, refmodel.NetG, refmodel.NetD = oV, oG, oD
    return net

# reg='Rec' against the reference's own dump of the same step
g = load_golden("rec_step")
net = build("Rec", 11, (4, 8), ([4, 4],))
assert torch.equal(net.net_mask.pruned, g["pruned"])
net.net_T.load_state_dict(sub(g, "sdT.")); net.net_R.load_state_dict(sub(g, "sdR."))
net.set_input(g["full"], g["aux"])
net.loss_all = 0
net.forwardT(); net.forwardR()
assert rel_l2(net.img_rec, g["img_rec"]) < 2e-5 and abs(net.loss_all.item() - g["loss_all"].item()) < 1e-5
before = net.net_R.cascades[1].dc_weight.detach().clone()
net.set_input(g["full"], g["aux"]); net.train(); net.update()
assert not torch.equal(before, net.net_R.cascades[1].dc_weight.detach())
# reg='Mixed': generator and discriminator passes of the reference's update()
g = load_golden("mixed_step")
net = build("Mixed", 12, (4, 8, 12, 8), ([4] * 2, [8] * 2, [8] * 2))
for t in "TRGD":
    getattr(net, "net_" + t).load_state_dict(sub(g, "sd" + t + "."))
net.set_input(g["full"], g["aux"])
net.loss_all = 0
net.forwardT(); net.forwardG(); net.forwardR(); net.forwardD(D_loss=False)
assert rel_l2(net.img_aligned, g["img_aligned"]) < 2e-4 and abs(net.loss_all.item() - g["loss_G"].item()) < 1e-4
snap = [p.detach().clone() for p in net.net_D.parameters()]
net.set_input(g["full"], g["aux"]); net.train(); net.update()
assert any(not torch.equal(a, b.detach()) for a, b in zip(snap, net.net_D.parameters()))
assert {"loss_gan_Dfake", "loss_gan_Dreal", "loss_gan_G", "loss_gan_sim"} <= set(net.get_vis("scalars")["scalars"])
# checkpoint through the reference's basemodel (one np.savez file per network): our modules' state_dicts save and
# load back through ckpt_save / ckpt_load
import os, tempfile
from basemodel import ckpt_load
d = os.path.join(tempfile.mkdtemp(), "ck")
net.save(d)
ck = ckpt_load(d)
assert {"net_G", "net_D", "net_T", "net_R", "net_mask", "config"} <= set(ck)
for t in "TRGD":
    getattr(net, "net_" + t).load_state_dict(ck["net_" + t])
print("dropin ok")
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_reference_model_py_drives_our_modules(tmp_path):
    script = tmp_path / "dropin.py"
    script.write_text(_SCRIPT)
    r = subprocess.run([sys.executable, str(script), ROOT, REF], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dropin ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
