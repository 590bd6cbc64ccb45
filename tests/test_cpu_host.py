"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol that
include/san_b200.h declares, the drop-in modules keep the reference's state_dict schema, the
product path refuses to run without CUDA (no fallback), masks are bit-exact, and the
data-parallel gradient exchange works at world_size 2 on gloo."""
import os
import random
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden

LIB = os.path.join(ROOT, "spatialalignmentnetwork_b200", "libsan_b200.so")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "build"], check=True)
    return LIB


def test_abi_symbols_exported(built):
    from spatialalignmentnetwork_b200 import _lib
    lib = _lib.lib()
    assert len(_lib.PROTOS) >= 40
    for name in _lib.PROTOS:
        assert hasattr(lib, name), name
    assert lib.san_version() >= 100
    assert lib.san_fft_workspace_bytes(2, 4, 8) == 2 * 4 * 8 * 8


def test_no_cpu_fallback(built):
    from spatialalignmentnetwork_b200 import signal_utils
    with pytest.raises(RuntimeError, match="CUDA"):
        signal_utils.fft2(torch.zeros(1, 1, 8, 8, dtype=torch.complex64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spatialalignmentnetwork_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_state_dict_schema_matches_golden():
    from spatialalignmentnetwork_b200.cross import SpatialTransformer
    from spatialalignmentnetwork_b200.varnet import VarNet
    g = load_golden("varnet_s")
    nc, ch, pools, sch, sp = [int(v) for v in g["cfg"]]
    net = VarNet(num_cascades=nc, sens_chans=sch, sens_pools=sp, chans=ch, pools=pools, use_ref=True)
    ref = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in sd)
    g = load_golden("align_s")
    ref = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    sd = SpatialTransformer(1).state_dict()
    assert set(sd.keys()) == set(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in sd)
    # full-size schema facts from the survey: 25 tensors per cascade, 24 in the sens net
    big = VarNet(num_cascades=8, sens_chans=8, sens_pools=4, chans=18, pools=4, use_ref=True).state_dict()
    assert len(big) == 224 and sum(v.numel() for v in big.values()) == 20120906
    assert big["cascades.0.model.unet.down_sample_layers.0.layers.0.weight"].shape == (18, 3, 3, 3)


def test_masks_bit_exact():
    from spatialalignmentnetwork_b200 import masks
    g = load_golden("masks")
    for shape, sp in ((320, 0.25), (320, 0.125), (368, 0.25), (64, 0.25)):
        random.seed(100 + shape)
        assert torch.equal(masks.EquispacedMask(sp, shape).pruned, g[f"equi_{shape}_{sp}"])
        torch.manual_seed(100 + shape)
        assert torch.equal(masks.StandardMask(sp, shape).pruned, g[f"std_{shape}_{sp}"])


def test_checkpoint_roundtrip(tmp_path):
    from spatialalignmentnetwork_b200 import model as M
    random.seed(3)
    cfg = M.Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg="Rec", mask="equispaced",
                   weight_smooth=1000.0, weight_sim=1.0, num_cascades=1)
    a = M.CSModel(cfg)
    a.save(str(tmp_path / "ckpt_1.pt"))
    b = M.CSModel(ckpt=str(tmp_path / "ckpt_1.pt"))
    for k, v in a.net_R.state_dict().items():
        assert torch.equal(v, b.net_R.state_dict()[k]), k
    assert torch.equal(a.net_mask.pruned, b.net_mask.pruned) and b.cfg.num_cascades == 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from spatialalignmentnetwork_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
torch.manual_seed(r)
lin = torch.nn.Linear(5, 3)
parallel.broadcast_state([lin])
w0 = lin.weight.detach().clone()
x = parallel.shard(torch.arange(20.).reshape(4, 5), r, 2)
lin(x).sum().backward()
local = lin.weight.grad.clone()
parallel.allreduce_mean_grads(lin.parameters())
both = [torch.zeros_like(local) for _ in range(2)]
dist.all_gather(both, local)
ws = [torch.zeros_like(w0) for _ in range(2)]
dist.all_gather(ws, w0)
assert torch.equal(ws[0], ws[1])
assert torch.allclose(lin.weight.grad, (both[0] + both[1]) / 2)
assert x.shape[0] == 2 and x[0, 0].item() == 10.0 * r
dist.destroy_process_group()
print("ok", r)
'''


def test_grad_allreduce_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + random.randint(0, 400))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_tc_geometry_covers_every_layer_shape(built):
    """Host-side geometry of the tcgen05 kernels (no GPU needed): every conv of the cascade / sensitivity /
    alignment U-Nets has a strip geometry (shared memory, TMEM columns, output-channel split) at the benchmark
    size, at the multi-coil 640x368 size and at small test sizes, for forward, data gradient and weight gradient."""
    from spatialalignmentnetwork_b200 import _lib
    L = _lib.lib()

    def unet_layers(cin, ch, pools):            # (Cin, Cout, K, level) of varnet.Unet
        out, c = [(cin, ch, 3, 0), (ch, ch, 3, 0)], ch
        for lv in range(1, pools):
            out += [(c, 2 * c, 3, lv), (2 * c, 2 * c, 3, lv)]
            c *= 2
        out += [(c, 2 * c, 3, pools), (2 * c, 2 * c, 3, pools)]
        for lv in range(pools - 1, -1, -1):
            out += [(2 * c, 4 * c, 1, lv + 1), (2 * c, c, 3, lv), (c, c, 3, lv)]     # convT as 1x1, cat conv, conv
            c //= 2
        return out + [(ch, 2, 1, 0)]

    align = [(2, 32, 3, 0), (32, 32, 3, 0), (96, 32, 3, 0), (32, 2, 3, 0)]
    for lv in range(1, 5):
        align += [(32 if lv == 1 else 64, 64, 1, lv), (64, 64, 3, lv), (128, 64, 3, lv), (64, 64, 1, lv)]
    for (H, W) in [(320, 320), (640, 368), (64, 64), (48, 32)]:
        for layers in (unet_layers(3, 18, 4), unet_layers(2, 8, 4), align):
            for cin, cout, k, lv in layers:
                h, w = H >> lv, W >> lv
                if h < 1 or w < 1:
                    continue
                assert L.san_tc_supported(h, w, cin, cout, k) == 1, (h, w, cin, cout, k)
                assert L.san_tc_supported(h, w, cout, cin, k) == 1, ("dgrad", h, w, cin, cout, k)
                assert L.san_tc_staged_weight_elems(h, w, cout, cin, k) > 0
                if w >= 16:
                    assert L.san_tc_wgrad_supported(h, w, cin, cout, k) == 1, ("wgrad", h, w, cin, cout, k)
    # staged activation buffer: [lead 8][N][2][Cpad/8][(H+2)(W+2)][8][trail 256]
    assert L.san_tc_staged_act_elems(2, 4, 6, 18) == 2 * 2 * 4 * 6 * 8 * 8 + 8 + 256
