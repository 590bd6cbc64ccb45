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

from conftest import ROOT, load_golden, rel_l2

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


def test_resume_cli_config_wins_and_missing_nets_raise(tmp_path):
    """Reference train.py:121 resumes with ``Model(ckpt=ckpt, cfg=cfg, objects=args.load_nets)``: the CLI config wins
    over the checkpoint's (the staged recipe of commands_train_test.sh changes --reg between stages), and a
    requested network that is not in the checkpoint is a KeyError (reference basemodel.py:178-181)."""
    import pytest
    from spatialalignmentnetwork_b200 import model as M
    random.seed(3)
    kw = dict(sparsity=0.25, lr=1e-4, shape=32, coils=1, mask="equispaced", weight_smooth=1000.0, weight_sim=1.0,
              num_cascades=1, gan_layers_G=[4, 8, 8], gan_layers_D=[[4, 4], [8, 8]])
    a = M.CSModel(M.Config(reg="None", **kw))
    a.save(str(tmp_path / "ckpt_1.pt"), objects=["net_mask", "net_R"])
    b = M.CSModel(ckpt=str(tmp_path / "ckpt_1.pt"), cfg=M.Config(reg="Mixed", **kw), objects=["net_mask", "net_R"])
    assert b.cfg.reg == "Mixed"
    assert all(torch.equal(v, b.net_R.state_dict()[k]) for k, v in a.net_R.state_dict().items())
    assert M.CSModel(ckpt=str(tmp_path / "ckpt_1.pt"), objects=["net_R"]).cfg.reg == "None"   # no cfg: the checkpoint's
    with pytest.raises(KeyError):
        M.CSModel(ckpt=str(tmp_path / "ckpt_1.pt"), cfg=M.Config(reg="Mixed", **kw), objects=["net_T"])
    with pytest.raises(KeyError):
        M.CSModel(ckpt=str(tmp_path / "ckpt_1.pt"), cfg=M.Config(reg="Mixed", **kw))      # all nets, but only 2 saved
    # train.py passes the CLI config on --resume and refuses --load_nets without it
    src = open(os.path.join(ROOT, "train.py")).read()
    assert "CSModel(ckpt=args.resume, cfg=cfg, objects=args.load_nets)" in src and "assert args.load_nets is None" in src


def test_mask_registry_and_lowpass_floor():
    """LowpassMask keeps floor(shape * sparsity) centre columns (reference masks.py:121); 'mask' is the plain
    fully-sampled Mask (reference model.py:30); the learned masks fail with an explicit message."""
    import pytest
    from spatialalignmentnetwork_b200 import masks as K
    m = K.masks["lowpass"](0.33, 320)
    assert int((~m.pruned).sum()) == 105                     # round() would give 106
    assert int(K.masks["mask"](320).pruned.sum()) == 0
    with pytest.raises(KeyError, match="unsupported mask"):
        K.masks["taylor"]


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
    # NetG(1, 1, (64, 128, 256, 512, 512)) / NetD(2, 5 blocks) of model.py:58-61 (gan.py:72-129): ConvDown as a
    # 1x1 conv over 4*Cin space-to-depth channels, concat convs, the 1-channel heads
    ch = [64, 128, 256, 512, 512]
    netg = [(1, 64, 3, 0), (64, 64, 3, 0), (192, 64, 3, 0), (64, 1, 3, 0)]
    for lv in range(1, 5):
        netg += [(ch[lv - 1] * 4, ch[lv], 1, lv), (ch[lv], ch[lv], 3, lv)]
        if lv < 4:
            netg += [(ch[lv] + ch[lv + 1], ch[lv], 3, lv)]
    netd = [(2, 64, 3, 0), (64, 64, 3, 0), (64, 128, 3, 1), (128, 128, 3, 1), (128, 256, 3, 2), (256, 256, 3, 2),
            (256, 256, 3, 3), (256, 256, 3, 4), (256, 1, 3, 4)]
    for (H, W) in [(320, 320), (640, 368), (64, 64), (48, 32)]:
        for layers in (unet_layers(3, 18, 4), unet_layers(2, 8, 4), align, netg, netd):
            for cin, cout, k, lv in layers:
                h, w = H >> lv, W >> lv
                if h < 1 or w < 1:
                    continue
                assert L.san_tc_supported(h, w, cin, cout, k) == 1, (h, w, cin, cout, k)
                assert L.san_tc_supported(h, w, cout, cin, k) == 1, ("dgrad", h, w, cin, cout, k)
                assert L.san_tc_staged_weight_elems(h, w, cout, cin, k) > 0
                if w >= 16:
                    assert L.san_tc_wgrad_supported(h, w, cin, cout, k) == 1, ("wgrad", h, w, cin, cout, k)
    # staged activation buffer: [lead 8][N][2][ceil(C/8)][(H+2)(W+2)][8][trail 256]: 18 channels = 3 groups
    assert L.san_tc_staged_act_elems(2, 4, 6, 18) == 2 * 2 * 3 * 6 * 8 * 8 + 8 + 256


def _describe(L, H, W, Cin, Cout, K):
    import ctypes
    out = (ctypes.c_int * 16)()
    assert L.san_tc_describe(H, W, Cin, Cout, K, ctypes.addressof(out)) == 0
    keys = ("Cin_pad KG KS nsplit Npad Wp Hp R T S_alloc strips stages acc_stages a_bytes b_bytes smem_bytes").split()
    g = dict(zip(keys, list(out)))
    form = (ctypes.c_int * 7)()
    assert L.san_tc_describe_form(H, W, Cin, Cout, K, ctypes.addressof(form)) == 0
    g.update(dict(zip("dxn Np wtaps xchg_bytes hls Ncol pair".split(), list(form))))
    return g


@pytest.mark.parametrize("shape", [(1, 3, 12, 20, 5, 3), (2, 18, 16, 24, 18, 3), (1, 20, 9, 33, 40, 1), (1, 7, 10, 18, 170, 3),
                                   (1, 36, 10, 160, 36, 3), (1, 32, 5, 320, 32, 3), (1, 2, 7, 50, 8, 3), (1, 72, 6, 40, 36, 3),
                                   (1, 40, 6, 24, 130, 3)])
def test_tc_flattened_pixel_formulation_on_cpu(built, shape):
    """CPU model of csrc/conv_tc.cu driven by the library's own geometry (san_tc_describe): the staged layout
    Xs[n][hl][kg][slot][8] with a one-pixel zero border, strips of R rows, 128-row M tiles over the flattened padded
    pixels q = r*Wp + x, filter taps as row offsets dy*Wp + dx of the same tile, BF16x3 partial products, output
    channel split -- reproduces conv2d.  Checks the index arithmetic and the resource bounds the kernel relies on,
    without a GPU."""
    import torch.nn.functional as F
    from spatialalignmentnetwork_b200 import _lib
    L = _lib.lib()
    N, Cin, H, W, Cout, K = shape
    g = _describe(L, H, W, Cin, Cout, K)
    Wp, Hp, R, T, Npad, KG = g["Wp"], g["Hp"], g["R"], g["T"], g["Npad"], g["KG"]
    ntaps = K * K
    # resource bounds
    assert g["Cin_pad"] % 16 == 0 and g["Cin_pad"] >= Cin and KG == -(-Cin // 8) and g["KS"] * 16 == g["Cin_pad"]
    assert Npad % 16 == 0 and g["nsplit"] * Npad >= Cout and Npad <= 256
    assert T * g["Ncol"] * g["acc_stages"] <= 512                              # TMEM columns
    hls_on = os.environ.get("SAN_TC_HLS", "1") != "0"
    assert g["hls"] == (1 if (hls_on and not g["dxn"] and K == 3 and Npad <= 32 and g["nsplit"] == 1) else 0)
    assert g["Ncol"] == (2 * Npad if g["hls"] else Npad)
    assert g["S_alloc"] >= 128 * T + 2 * Wp + 2 and g["S_alloc"] >= (R + 2) * Wp  # every tap row of every tile is in the tile
    dxn, Np = g["dxn"], g["Np"]
    dxn_on = os.environ.get("SAN_TC_DXN", "0") != "0"
    assert dxn == (1 if (dxn_on and K == 3 and Cout <= 40) else 0)   # narrow 3x3 layers: horizontal taps in the MMA N dimension
    assert g["wtaps"] == (3 if dxn else ntaps)
    assert g["a_bytes"] == 4 * g["S_alloc"] * 16 and g["b_bytes"] == g["wtaps"] * 4 * Npad * 16
    assert g["stages"] >= 2 and 256 + g["xchg_bytes"] + g["stages"] * (g["a_bytes"] + g["b_bytes"]) <= 225 * 1024
    assert g["strips"] == -(-H // R) and 128 * T >= R * Wp
    if dxn:
        assert Np % 8 == 0 and Np >= Cout and Npad >= 3 * Np and g["nsplit"] == 1
        assert g["xchg_bytes"] >= T * 4 * 5 * Np * 4                   # [T][quarter][5][Np] floats
    torch.manual_seed(7)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, K, K) / (Cin * ntaps) ** 0.5
    ref = F.conv2d(x.double(), w.double(), padding=K // 2)

    def split(t):           # BF16 hi / lo halves
        hi = t.to(torch.bfloat16).float()
        return hi, (t - hi).to(torch.bfloat16).float()

    # staged activations [N][hl][KG][Hp*Wp][8] (zero border, zero pad channels; KG = ceil(Cin / 8) REAL groups: with an
    # odd count the padding group of the last K-step has no staged plane - modelled as a zero plane in the tile below,
    # standing for "never read" (tap pairing) or "finite stale shared memory x zero weight rows" (otherwise)
    KGp = g["Cin_pad"] // 8
    xp = torch.zeros(N, g["Cin_pad"], Hp, Wp)
    xp[:, :Cin, 1:H + 1, 1:W + 1] = x
    xs = torch.stack(split(xp), 1).reshape(N, 2, KGp, 8, Hp * Wp).permute(0, 1, 2, 4, 3)     # [N,hl,KGp,slot,8]
    assert float(xs[:, :, KG:].abs().max() if KGp > KG else 0.0) == 0.0                      # what is not staged is all zero
    wp_ = torch.zeros(g["nsplit"] * Npad, g["Cin_pad"], ntaps)
    wp_[:Cout, :Cin] = w.reshape(Cout, Cin, ntaps)
    whi, wlo = split(wp_)
    out = torch.zeros(N, Cout, H, W, dtype=torch.float64)
    for n in range(N):
        for st in range(g["strips"]):
            y0 = st * R
            rows_in = min(R + 2, Hp - y0)
            tile = torch.zeros(2, KGp, g["S_alloc"], 8)                        # what the TMA bulk copies deliver
            tile[:, :, :rows_in * Wp] = xs[n][:, :, y0 * Wp:(y0 + rows_in) * Wp]
            if dxn:
                # B rows nn = dx * Np + co per filter row dy; ONE A window per dy (offset dy * Wp); the accumulator
                # holds E_dx[q] in column block dx, and y[q] = E_0[q] + E_1[q + 1] + E_2[q + 2]
                acc = torch.zeros(128 * T, Npad, dtype=torch.float64)
                for dy in range(3):
                    bh = torch.zeros(Npad, g["Cin_pad"], dtype=torch.float64)
                    bl = torch.zeros(Npad, g["Cin_pad"], dtype=torch.float64)
                    for dx in range(3):
                        bh[dx * Np:dx * Np + Cout] = whi[:Cout, :, dy * 3 + dx].double()
                        bl[dx * Np:dx * Np + Cout] = wlo[:Cout, :, dy * 3 + dx].double()
                    off = dy * Wp
                    assert off + 128 * T <= g["S_alloc"]
                    a_hi = tile[0, :, off:off + 128 * T].permute(1, 0, 2).reshape(128 * T, -1).double()
                    a_lo = tile[1, :, off:off + 128 * T].permute(1, 0, 2).reshape(128 * T, -1).double()
                    acc += a_hi @ bh.T + a_lo @ bh.T + a_hi @ bl.T
                for q in range(R * Wp):
                    r, xx = divmod(q, Wp)
                    if xx < W and y0 + r < H:
                        assert q + 2 < 128 * T          # the neighbouring lanes exist (next quarter / next tile of the unit)
                        out[n, :, y0 + r, xx] = (acc[q, 0:Cout] + acc[q + 1, Np:Np + Cout]) + acc[q + 2, 2 * Np:2 * Np + Cout]
                continue
            pair_on = os.environ.get("SAN_TC_PAIR", "1") != "0"
            assert g["pair"] == (1 if (pair_on and K == 3 and not dxn and (-(-Cin // 8)) % 2 == 1) else 0)
            for ns in range(g["nsplit"]):
                acc = torch.zeros(128 * T, Npad, dtype=torch.float64)
                acc2 = torch.zeros(128 * T, 2 * Npad, dtype=torch.float64)
                offs = [(tap // 3) * Wp + (tap % 3) if ntaps == 9 else Wp + 1 for tap in range(ntaps)]

                def mma(a_h, a_l, b_h, b_l):            # one MMA step of K = 16 (operands [rows, 16] / [Npad, 16])
                    nonlocal acc, acc2
                    if g["hls"]:     # B = [W_hi | W_lo] stacked along N: A_hi x B (2 Npad columns) + A_lo x W_hi (first Npad columns)
                        acc2 += a_h @ torch.cat([b_h, b_l], 0).T
                        acc2[:, :Npad] += a_l @ b_h.T
                    else:
                        acc += a_h @ b_h.T + a_l @ b_h.T + a_h @ b_l.T

                def window(hl, kg, off):                 # [128 T, 8]: what the descriptor reads for one K group
                    assert off + 128 * T <= g["S_alloc"]
                    return tile[hl, kg, off:off + 128 * T].double()

                for ks in range(g["KS"]):
                    wsl = lambda t_, lo: (wlo if lo else whi)[ns * Npad:(ns + 1) * Npad, ks * 16:ks * 16 + 16, t_].double()
                    if g["pair"] and ks == g["KS"] - 1:
                        # one real channel group (2 ks): K group 0 = its window at tap 2i, K group 1 = the SAME plane at
                        # tap 2i + 1 (descriptor LBO = window distance); tap 8 pairs with itself against zero weights.
                        # The all-zero group 2 ks + 1 is never read.
                        for i in range(5):
                            t0, t1 = 2 * i, min(2 * i + 1, 8)
                            lbo = offs[t1] - offs[t0] if i < 4 else 0
                            assert 0 <= lbo < 16384
                            a = [torch.cat([window(hl, 2 * ks, offs[t0]), window(hl, 2 * ks, offs[t0] + lbo)], 1) for hl in (0, 1)]
                            z = torch.zeros(Npad, 8, dtype=torch.float64)
                            b = [torch.cat([wsl(t0, lo)[:, :8], wsl(t1, lo)[:, :8] if i < 4 else z], 1) for lo in (0, 1)]
                            assert float(wsl(t0, 0)[:, 8:].abs().max()) == 0      # the skipped group really is all zero
                            mma(a[0], a[1], b[0], b[1])
                    else:
                        for tap in range(ntaps):
                            a = [torch.cat([window(hl, 2 * ks, offs[tap]), window(hl, 2 * ks + 1, offs[tap])], 1) for hl in (0, 1)]
                            mma(a[0], a[1], wsl(tap, 0), wsl(tap, 1))
                if g["hls"]:
                    acc = acc2[:, :Npad] + acc2[:, Npad:]          # epilogue: the two column blocks of a pixel
                for q in range(R * Wp):
                    r, xx = divmod(q, Wp)
                    if xx < W and y0 + r < H:
                        c_cnt = min(Npad, Cout - ns * Npad)
                        if c_cnt > 0:
                            out[n, ns * Npad:ns * Npad + c_cnt, y0 + r, xx] = acc[q, :c_cnt]
    assert rel_l2(out, ref) < 2e-5


@pytest.mark.parametrize("shape", [(3, 18, 7, 36, 18), (2, 3, 5, 40, 18), (2, 16, 6, 33, 8), (1, 24, 4, 130, 32), (2, 18, 5, 320, 18)])
@pytest.mark.parametrize("nctas", [1, 3])
def test_tc_row_ring_formulation_on_cpu(built, shape, nctas):
    """CPU model of conv_rows_kernel (csrc/conv_tc.cu: the conv that stages its own input) driven by the library's geometry
    (san_tc_conv_rows_describe): contiguous unit ranges per CTA, the converted padded rows in a ring of NR shared-memory row
    slots [hl][kg][RS][8], one-row units whose tap windows lie inside single ring rows, HLS accumulators, taps of the
    half-empty last K-step paired INSIDE a filter row (pair = 2 weight layout) -- reproduces conv2d; the ring protocol
    (a slot is refilled only after the unit that last read it, every row is freed exactly once, no unit waits for a row
    whose slot depends on itself) and the resource bounds are checked on the way.  No GPU."""
    import ctypes
    import torch.nn.functional as F
    from spatialalignmentnetwork_b200 import _lib
    L = _lib.lib()
    N, Cin, H, W, Cout = shape
    out_g = (ctypes.c_int * 12)()
    assert L.san_tc_conv_rows_describe(H, W, Cin, Cout, 1, ctypes.addressof(out_g)) == 0
    KG, KS, Npad, Ncol, Wp, T, RS, pair, NR, row_bytes, w_bytes, smem = list(out_g)
    assert KG == -(-Cin // 8) <= 3 and KS == (KG + 1) // 2 and pair == (2 if KG % 2 else 0)
    assert Npad % 16 == 0 and Cout <= Npad <= 32 and Ncol == 2 * Npad and 2 * T * Ncol <= 512      # TMEM, double-buffered
    assert Wp == W + 2 and 128 * T >= Wp and RS >= 128 * (T - 1) + 2 + 1 + 128                      # every window inside the plane
    assert row_bytes == 2 * KG * RS * 16 and w_bytes == KS * 9 * 4 * Npad * 16
    assert 4 <= NR <= 8 and smem <= 225 * 1024 and 256 + w_bytes + NR * row_bytes <= smem
    assert Wp * KG <= 4 * 256                                                                    # converter items per group and row
    torch.manual_seed(11)
    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, 3, 3) / (Cin * 9) ** 0.5
    ref = F.conv2d(x.double(), w.double(), padding=1)

    def split(t):
        hi = t.to(torch.bfloat16).float()
        return hi, (t - hi).to(torch.bfloat16).float()

    # weight blocks as stage_weights_kernel writes them (hls = 1, pair as described): blocks[ks][blk] = (B_hi, B_lo) [Npad, 16]
    wp_ = torch.zeros(Npad, 16 * KS, 9)
    wp_[:Cout, :Cin] = w.reshape(Cout, Cin, 9)
    whi, wlo = split(wp_)
    blocks = []
    for ks in range(KS):
        row = []
        if pair and ks == KS - 1:
            for blk in range(6):
                dy, odd = blk >> 1, blk & 1
                t0 = dy * 3 + (2 if odd else 0)
                z = torch.zeros(Npad, 8)
                bh = torch.cat([whi[:, ks * 16:ks * 16 + 8, t0], z if odd else whi[:, ks * 16:ks * 16 + 8, dy * 3 + 1]], 1)
                bl = torch.cat([wlo[:, ks * 16:ks * 16 + 8, t0], z if odd else wlo[:, ks * 16:ks * 16 + 8, dy * 3 + 1]], 1)
                row.append((bh.double(), bl.double(), dy, 2 if odd else 0, 0 if odd else 1))     # (.., dy, dx of K group 0, LBO in slots)
            assert float(whi[:, ks * 16 + 8:ks * 16 + 16].abs().max()) == 0                      # the unstaged padding group
        else:
            for tap in range(9):
                row.append((whi[:, ks * 16:ks * 16 + 16, tap].double(), wlo[:, ks * 16:ks * 16 + 16, tap].double(), tap // 3, tap % 3, None))
        blocks.append(row)

    def convert(n, pr):          # one padded row as the converters write it: [hl][kg][RS][8], zeros outside the image / past Wp
        rowt = torch.zeros(2, KG, RS, 8)
        if 1 <= pr <= H:
            v = torch.zeros(KG * 8, Wp)
            v[:Cin, 1:W + 1] = x[n, :, pr - 1]
            hi, lo = split(v)
            rowt[0, :, :Wp] = hi.reshape(KG, 8, Wp).permute(0, 2, 1)
            rowt[1, :, :Wp] = lo.reshape(KG, 8, Wp).permute(0, 2, 1)
        return rowt

    out = torch.zeros(N, Cout, H, W, dtype=torch.float64)
    nunits = N * H
    for cta in range(nctas):
        u0, u1 = cta * nunits // nctas, (cta + 1) * nunits // nctas
        ring = [None] * NR                  # ring[s] = (row counter, tensor)
        freed = set()
        k_next = 0
        for u in range(u0, u1):
            n, yy = divmod(u, H)
            fresh = u == u0 or yy == 0
            base = k_next if fresh else k_next - 2
            for r in (range(3) if fresh else range(2, 3)):       # converters: new rows of this unit, in counter order
                k = k_next
                s_ = k % NR
                if ring[s_] is not None:
                    assert ring[s_][0] == k - NR and ring[s_][0] in freed, "slot refilled before its row was released"
                ring[s_] = (k, convert(n, yy + r))
                k_next += 1
            rows = []
            for dy in range(3):
                kk_, t_ = ring[(base + dy) % NR]
                assert kk_ == base + dy                           # the MMA warp finds the unit's rows where it expects them
                rows.append(t_)
            for t in range(T):
                acc2 = torch.zeros(128, 2 * Npad, dtype=torch.float64)
                for ks in range(KS):
                    for (bh, bl, dy, dx, lbo) in blocks[ks]:
                        s0 = t * 128 + dx
                        if lbo is None:                           # K group 1 = the next channel-group plane
                            kg0 = 2 * ks
                            a = [torch.cat([rows[dy][hl, kg0, s0:s0 + 128], rows[dy][hl, kg0 + 1, s0:s0 + 128]], 1).double()
                                 for hl in (0, 1)]
                        else:                                     # K group 1 = the same plane, lbo slots further (0: against zero weights)
                            assert s0 + lbo + 128 <= RS
                            a = [torch.cat([rows[dy][hl, KG - 1, s0:s0 + 128], rows[dy][hl, KG - 1, s0 + lbo:s0 + lbo + 128]], 1).double()
                                 for hl in (0, 1)]
                        acc2 += a[0] @ torch.cat([bh, bl], 0).T
                        acc2[:, :Npad] += a[1] @ bh.T
                acc = acc2[:, :Npad] + acc2[:, Npad:]
                for lane in range(128):
                    xx = t * 128 + lane
                    if xx < W:
                        out[n, :, yy, xx] = acc[lane, :Cout]
            # rows no later unit reads
            next_fresh = u + 1 == u1 or (u + 1) % H == 0
            for kf in ([base, base + 1, base + 2] if next_fresh else [base]):
                assert kf not in freed
                freed.add(kf)
        assert freed == set(range(k_next))                        # every converted row was released exactly once
    assert rel_l2(out, ref) < 2e-5


@pytest.mark.parametrize("shape", [(2, 3, 12, 20, 5, 3), (1, 18, 16, 24, 18, 3), (1, 20, 9, 33, 40, 1), (1, 7, 10, 18, 170, 3),
                                   (1, 192, 6, 16, 24, 3), (1, 2048, 3, 16, 16, 1)])
def test_tc_wgrad_formulation_on_cpu(built, shape):
    """CPU model of csrc/wgrad_tc.cu driven by san_tc_wgrad_describe: groups (128-channel block of dY, chunk of
    input channels, filter row), pixel chunks of KC slots over the range [Wp, Wp + ceil16(H*Wp)) of every image,
    K = 16-slot GEMM steps, the X span starting (dy-1)*Wp - 1 slots before the dY chunk with dx as a +1-slot
    shift, zero border pixels of dY instead of masks, lead-in / trailing slack of the staged buffer -- reproduces
    the conv2d weight gradient."""
    import ctypes
    import torch.nn.functional as F
    from spatialalignmentnetwork_b200 import _lib
    L = _lib.lib()
    N, Cin, H, W, Cout, K = shape
    out = (ctypes.c_int * 16)()
    assert L.san_tc_wgrad_describe(H, W, Cin, Cout, K, ctypes.addressof(out)) == 0
    g = dict(zip("KGo KGi nmb nnc ndy Nn KGn KC XS stages smem_bytes nchunks Wp PS range0 range_len".split(), list(out)))
    Wp, PS, KC = g["Wp"], g["PS"], g["KC"]
    import os
    assert g["KGo"] == -(-Cout // 8) and g["KGi"] == -(-Cin // 8)            # real channel groups = staged planes
    narrow = g["nnc"] == 1 and g["nmb"] == 1 and 2 * g["KGo"] <= 8           # M = 64 shape: N = the real groups only
    assert g["Nn"] == (8 * g["KGi"] if narrow else (-(-Cin // 16) * 16) // g["nnc"]) and g["KGn"] * 8 == g["Nn"]
    assert g["Nn"] % (8 if narrow else 16) == 0 and g["Nn"] <= 160 and 3 * g["Nn"] <= 512
    assert g["nmb"] == -(-g["KGo"] // 16) and KC % 16 == 0 and g["range_len"] % 16 == 0 and g["stages"] >= 2
    assert g["range0"] == Wp and g["range0"] + g["range_len"] <= PS and g["smem_bytes"] <= 225 * 1024
    assert g["nchunks"] == -(-g["range_len"] // KC)
    torch.manual_seed(9)
    x = torch.randn(N, Cin, H, W)
    gy = torch.randn(N, Cout, H, W)
    wr = torch.zeros(Cout, Cin, K, K, dtype=torch.float64, requires_grad=True)
    (F.conv2d(x.double(), wr, padding=K // 2) * gy.double()).sum().backward()
    LEAD, TRAIL = 8, 256

    def staged(t, C):       # flat buffer [lead][N][hl][KG][PS][8][trail] of hi/lo values (as float), KG = ceil(C / 8)
        Cp = (C + 7) // 8 * 8
        tp = torch.zeros(N, Cp, H + 2, W + 2)
        tp[:, :C, 1:H + 1, 1:W + 1] = t
        hi = tp.to(torch.bfloat16).float()
        lo = (tp - hi).to(torch.bfloat16).float()
        body = torch.stack([hi, lo], 1).reshape(N, 2, Cp // 8, 8, PS).permute(0, 1, 2, 4, 3).reshape(-1)
        return torch.cat([torch.zeros(LEAD), body, torch.zeros(TRAIL)]), Cp // 8

    dys, KGo = staged(gy, Cout)
    xs, KGi = staged(x, Cin)
    assert (KGo, KGi) == (g["KGo"], g["KGi"])

    def span(buf, KGtot, n, hl, kg, slot0, nslots):      # what one bulk copy delivers: [nslots, 8]
        o = LEAD + (((n * 2 + hl) * KGtot + kg) * PS + slot0) * 8
        assert o >= 0 and o + nslots * 8 <= buf.numel()
        return buf[o:o + nslots * 8].reshape(nslots, 8).double()

    form = (ctypes.c_int * 2)()
    assert L.san_tc_wgrad_describe_form(H, W, Cin, Cout, K, ctypes.addressof(form)) == 0
    rown, ncp = form[0], form[1]
    # "filter rows in N": 3x3 layers with <= 48 padded input channels load three row-shifted copies of the X span
    # and fold the filter rows into the MMA N dimension (one CTA group accumulates all nine taps)
    assert rown == (1 if (K == 3 and 9 * g["Nn"] <= 512 and os.environ.get("SAN_WG_ROWN", "1") != "0") else 0)
    assert ncp == (3 if rown else 1) and g["ndy"] == (1 if rown else K)
    ndx = K
    dw = torch.zeros(Cout, Cin, K, K, dtype=torch.float64)
    for mb in range(g["nmb"]):
        kga = min(16, KGo - 16 * mb)
        for nc in range(g["nnc"]):
            for dyi in range(g["ndy"]):
                rows = list(range(3)) if rown else [dyi]        # filter rows covered by this group's MMAs
                Nmma = len(rows) * g["Nn"]
                assert ndx * Nmma <= 512                         # TMEM columns of the dx accumulators
                D = [torch.zeros(kga * 8, Nmma, dtype=torch.float64) for _ in range(ndx)]
                for n in range(N):
                    for ch in range(g["nchunks"]):
                        p0 = g["range0"] + ch * KC
                        kc = min(KC, g["range_len"] - ch * KC)
                        A = [torch.cat([span(dys, KGo, n, hl, 16 * mb + k, p0, kc) for k in range(kga)], 1) for hl in (0, 1)]
                        # B planes [hl][row copy][kg]: the N index of the MMA is (row copy, channel)
                        kgl = min(g["KGn"], KGi - nc * g["KGn"])        # real groups of this chunk; a padding group is never
                        zpl = torch.zeros(kc + 2, 8, dtype=torch.float64)   # loaded (stale plane, dW columns dropped): zeros here
                        B = [torch.cat([span(xs, KGi, n, hl, nc * g["KGn"] + k, p0 + ((dy - 1) * Wp - 1 if K == 3 else 0), kc + 2)
                                        if k < kgl else zpl for dy in rows for k in range(g["KGn"])], 1) for hl in (0, 1)]
                        for dx in range(ndx):
                            Bh, Bl = B[0][dx:dx + kc], B[1][dx:dx + kc]
                            D[dx] += A[0].T @ Bh + A[1].T @ Bh + A[0].T @ Bl
                for dx in range(ndx):
                    for ri, dy in enumerate(rows):
                        tap = (dy, dx) if K == 3 else (0, 0)
                        co0, ci0 = mb * 128, nc * g["Nn"]
                        nco, nci = min(kga * 8, Cout - co0), min(g["Nn"], Cin - ci0)
                        if nco > 0 and nci > 0:
                            dw[co0:co0 + nco, ci0:ci0 + nci, tap[0], tap[1]] += D[dx][:nco, ri * g["Nn"]:ri * g["Nn"] + nci]
    assert rel_l2(dw, wr.grad) < 2e-5


def test_fft_small_host(tmp_path):
    """csrc/fft_small.cuh (the per-thread register DFTs and the two-phase 16 x 20 decomposition of the opt-in v2 FFT
    kernels, SAN_FFT_V2=1) compiled for the HOST and checked against a direct fp64 DFT."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft_small_test")
    src = os.path.join(ROOT, "tests", "host", "fft_small_test.cu")
    b = subprocess.run([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                       capture_output=True, text=True, timeout=300)
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr


_WORKER_ATTACH = r'''
import os, random, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from spatialalignmentnetwork_b200 import model as M, parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
random.seed(7)                      # identical column masks on every rank (train.py seeds python random the same way)
torch.manual_seed(100 + r)          # different initial weights: attach() must make them identical
cfg = M.Config(sparsity=0.25, lr=1e-4, shape=32, coils=1, reg="Mixed", mask="equispaced", weight_smooth=1000.0,
               weight_sim=1.0, weight_gan=0.1, weight_gan_sim=1.0, num_cascades=1, fused_adamw=False,
               gan_layers_G=[4, 8, 8], gan_layers_D=[[4, 4], [8, 8]])
net = M.CSModel(cfg)
parallel.attach(net)
for name in ("net_mask", "net_G", "net_D", "net_T", "net_R"):
    for k, v in getattr(net, name).state_dict().items():
        both = [torch.zeros_like(v) for _ in range(2)]
        dist.all_gather(both, v.contiguous())
        assert torch.equal(both[0], both[1]), (name, k)
# the bracket CSModel.update() puts around backward(): _arm(nets) ... _sync(nets) = mean over ranks of exactly those
# networks (T, G, R in the generator step, D in the discriminator step)
net._arm([net.net_T, net.net_G])
for p in net.net_G.parameters():
    p.grad.fill_(float(r + 1))          # .grad is a view of the flat bucket buffer
for p in net.net_D.parameters():
    p.grad = torch.full_like(p, 10.0 * (r + 1))
net._sync([net.net_T, net.net_G])   # net_T received no gradients: zeros
assert all(torch.all(p.grad == 1.5) for p in net.net_G.parameters())
assert all(torch.all(p.grad == 0.0) for p in net.net_T.parameters())
assert all(torch.all(p.grad == 10.0 * (r + 1)) for p in net.net_D.parameters())
net._arm([net.net_D])
for p in net.net_D.parameters():
    p.grad.fill_(10.0 * (r + 1))
net._sync([net.net_D])
assert all(torch.all(p.grad == 15.0) for p in net.net_D.parameters())
# blocking fallback (attach(overlap=False)): same result through grad_sync
net.grad_buckets = None
for p in net.net_G.parameters():
    p.grad = torch.full_like(p, float(r + 1))
net._sync([net.net_G])
assert all(torch.all(p.grad == 1.5) for p in net.net_G.parameters())
dist.destroy_process_group()
print("ok", r)
'''


_WORKER_BUCKETS = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from spatialalignmentnetwork_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()

class R(torch.nn.Module):            # parameter names like the VarNet's: sens_net.*, cascades.<i>.*
    def __init__(self):
        super().__init__()
        self.sens_net = torch.nn.Linear(6, 6)
        self.cascades = torch.nn.ModuleList([torch.nn.Linear(6, 6) for _ in range(3)])
        self.unused = torch.nn.Parameter(torch.ones(2))       # receives no gradient: its bucket leaves in sync()
    def forward(self, x):
        x = self.sens_net(x)
        for c in self.cascades:
            x = torch.tanh(c(x))
        return x

class Model:
    pass

torch.manual_seed(5 + r)
m = Model()
m.net_R, m.net_T = R(), torch.nn.Linear(6, 6)
parallel.attach(m)
gb = m.grad_buckets
assert [b.key for b in gb.buckets["net_R"]] == ["unused", "sens_net", "cascades.0", "cascades.1", "cascades.2"]
assert [b.key for b in gb.buckets["net_T"]] == ["net_T"]
x = parallel.shard(torch.arange(24.).reshape(4, 6) / 10, r, 2)
for step in range(2):                # second step: buffers are re-zeroed, aliasing survives zero_grad(set_to_none)
    for p in list(m.net_R.parameters()) + list(m.net_T.parameters()):
        p.grad = None
    gb.arm(["net_T", "net_R"])
    n0 = gb.launched
    loss = m.net_R(m.net_T(x)).pow(2).sum()
    loss.backward()
    assert gb.launched - n0 == 5, gb.launched - n0   # net_T + sens + 3 cascades left from the hooks (overlapped)
    gb.sync(["net_T", "net_R"])
    got = [p.grad.clone() for p in list(m.net_T.parameters()) + list(m.net_R.parameters())]
    for p, v in zip(gb.params["net_R"], gb.views["net_R"]):
        assert p.grad.data_ptr() == v.data_ptr()             # .grad is a view of the flat buffer: no copies
    # reference: plain local gradients, blocking all-reduce
    for p in list(m.net_R.parameters()) + list(m.net_T.parameters()):
        p.grad = None
    m.net_R(m.net_T(x)).pow(2).sum().backward()
    m.net_R.unused.grad = torch.zeros(2)
    ps = list(m.net_T.parameters()) + list(m.net_R.parameters())
    parallel.allreduce_mean_grads(ps)
    for a, p in zip(got, ps):
        assert torch.allclose(a, p.grad, rtol=1e-6, atol=1e-7)
# a network that is not armed is left alone (Mixed mode: net_D collects gradients in the generator step)
for p in m.net_T.parameters():
    p.grad = None
gb.arm(["net_R"])
m.net_R(m.net_T(x)).sum().backward()
gb.sync(["net_R"])
both = [torch.zeros_like(m.net_T.weight.grad) for _ in range(2)]
dist.all_gather(both, m.net_T.weight.grad)
assert not torch.equal(both[0], both[1])
dist.destroy_process_group()
print("ok", r)
"""


def test_grad_buckets_overlapped_world_size_2(tmp_path):
    """parallel.GradBuckets: per-cascade buckets in one flat buffer per network, launched from autograd hooks during
    backward, equal to the blocking flat all-reduce; .grad aliases the flat buffer."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER_BUCKETS)
    port = str(29850 + random.randint(0, 40))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_attach_broadcasts_all_networks_world_size_2(tmp_path):
    """parallel.attach on a CSModel with the GAN networks: every parameter / buffer (spectral-norm u, v and
    BatchNorm statistics included) identical on both ranks afterwards; the gradient hook averages exactly the
    networks ``update()`` names (T, G, R in the generator step, D in the discriminator step)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER_ATTACH)
    port = str(29950 + random.randint(0, 40))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_conv_cycle_model_runs(built):
    """tools/conv_model.py (analytic MMA / operand-read / HBM model over the kernel's own geometry) stays runnable
    and keeps explaining the measured per-layer times of profiles/r1h_conv_tc_microbench.txt within 2.2x."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import conv_model
    from spatialalignmentnetwork_b200 import _lib
    L = _lib.lib()
    for shape, meas in conv_model.MEASURED_MS.items():
        if shape == (18, 2, 320, 1):       # 1x1 head: the model charges halo rows the 1x1 kernel finds in L2
            continue
        g, m = conv_model.model(L, 64, *shape)
        if (g["dxn"], g["hls"]) != conv_model.MEASURED_FORM.get(shape, (0, 0)):
            continue                       # measured with another formulation
        assert 0.95 < meas / m["bound"] < 2.2, (shape, meas, m)


def test_fft_v2_kernels_under_host_emulation(tmp_path):
    """The opt-in v2 FFT kernels (csrc/fft_v2.cuh, SAN_FFT_V2=1), compiled from the SAME source for the host with one
    OS thread per CUDA thread and a barrier for __syncthreads(): every fused load / store variant (plain, ACS column
    mask, planar, sens expand + soft DC, coil-reducing stores with 1 and 2 coils, a ragged column CTA), forward and
    inverse, against a direct fp64 2-D DFT at 320x320."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    assert cxx, "g++ not found"
    exe = str(tmp_path / "fft_v2_emul")
    src = os.path.join(ROOT, "tests", "host", "fft_v2_emul.cpp")
    b = subprocess.run([cxx, "-std=c++17", "-O2", "-pthread", "-I/usr/local/cuda/include", "-DSAN_FFT_EMULATE", "-o", exe, src],
                       capture_output=True, text=True, timeout=300)
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "worst relative error" in r.stdout


def test_abi_argument_validation_without_gpu(built):
    """Error behaviour of the C ABI (reference: bare asserts in Python; here return codes + san_last_error): bad
    arguments are rejected before any CUDA call, so this runs without a GPU."""
    from spatialalignmentnetwork_b200 import _lib
    L = _lib.lib()
    ARG, UNSUP = -1, -3
    p = 0x1000      # a non-null "device pointer" that is never dereferenced: validation fails first
    cases = [
        (L.san_fft2(None, 0, None, p, 0, None, p, 1, 8, 8, 0, None), "null"),
        (L.san_fft2(p, 1, p, p, 0, None, p, 1, 8, 8, 0, None), "column mask"),
        (L.san_fft_expand_dc(p, p, p, None, None, None, p, p, 1, 1, 8, 8, 0, None), "k0"),
        (L.san_pair_loss_fwd(p, None, 16, 0, 1.0, p, p, None), "bad args"),          # L1 without y
        (L.san_pair_loss_fwd(p, p, 16, 3, 1.0, p, p, None), "bad args"),             # unknown mode
        (L.san_mi_metric(p, p, 1, 64, 65, 0.0, 1.0, p, None), "bins"),
        (L.san_mi_metric(p, p, 1, 64, 64, 1.0, 1.0, p, None), "maxv"),
        (L.san_warp_reflect(p, p, p, 1, 1, 8, 8, 8, 8, 3, None), "bad args"),
        (L.san_filter2d(p, p, p, 1, 8, 8, 4, None), "bad args"),                     # even window
        (L.san_sn_sigma(p, p, p, p, p, 0, 4, 1e-12, 1, None), "bad args"),
        (L.san_adamw_step(p, p, p, p, p, 1, 1e-4, 0.9, 0.999, 1e-8, 0.0, 0, None, None, None), "bad args"),     # step counts from 1
        (L.san_tc_conv(p, p, None, p, 1, 8, 8, 4, 4, 5, 0, 3, None, None), "unsupported shape"),  # 5x5 filter
        (L.san_tc_conv(p, p, None, p, 1, 8, 8, 4, 4, 3, 0, 2, None, None), "bad args"),     # mixed pair formats fault on the B200
        (L.san_tc_conv(p, p, None, p, 1, 8, 8, 4, 4, 3, 0, 0, p, None), "dynamic scale"),  # dynamic scale needs fp16 pairs
        (L.san_tc_stage_terms(p, 1, 8, 8, 16, p, 7, 1, None, None), "bad args"),     # more than 6 terms
    ]
    for rc, msg in cases:
        assert rc == ARG, (rc, msg)
    import ctypes
    out = (ctypes.c_int * 16)()
    assert L.san_tc_describe(8, 8, 4, 4, 5, ctypes.addressof(out)) == UNSUP
    assert L.san_tc_wgrad_describe(8, 8, 4, 4, 3, ctypes.addressof(out)) == UNSUP       # image narrower than 16 pixels
    # the message of the LAST failure on this thread
    assert L.san_mi_metric(p, p, 1, 64, 65, 0.0, 1.0, p, None) == ARG and b"bins" in L.san_last_error()


def test_tap_pairing_design_study():
    """tools/tap_pairing_model.py: the proposed K-padding-free formulation for the 18- / 36-channel layers (leftover
    channel group of two filter taps in one K = 16 UMMA step, via the descriptor's leading-dimension offset) checked at
    the level of descriptor arithmetic against conv2d."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import tap_pairing_model
    tap_pairing_model.main()
