#!/usr/bin/env python
"""Benchmark of the VarNet + spatial-alignment hot path (BASELINE.json metric: slices/sec).

One "step" = what ``train.py`` does per iteration for ``--reg Rec`` (reference train.py:207-217,
model.py:89-121, 142-169, 206-216) on one batch of synthetic 320x320 complex64 T1/T2 slice pairs:
``set_input`` (fft2, mask, ifft2, rss) -> ``forwardT`` (alignment U-Net, displacement field, bilinear
warp, smoothness loss) -> ``forwardR`` (12-cascade VarNet with the warped reference channel, SSIM
loss) -> backward -> gradient all-reduce (N > 1) -> AdamW step of ``net_T`` and ``net_R``.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # CPU oracle port of the reference

Prints ONE JSON line (rank 0).  ``value`` = slices/s with the inputs resident in HBM;
``e2e`` = the same step driven from pinned HOST buffers (H2D of both modalities + D2H of the loss
inside the timed region); ``roofline`` = the dominant kernel class, timed per launch with CUDA
events on the launching stream in one extra instrumented step; ``cpu_baseline`` = the CPU oracle
(port of the reference's PyTorch path) on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20221017
METRIC = "slices/sec (320x320 complex, 12-cascade VarNet+align, fwd+bwd+optimizer step)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="slices per GPU per step (weak scaling)")
    ap.add_argument("--shape", default="320", help="H (square) or HxW, e.g. 640x368 (BASELINE cfg4)")
    ap.add_argument("--coils", type=int, default=1, help="coils per slice (cfg4: 15)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch slices per GPU; strong: --batch is the GLOBAL batch, split over the GPUs")
    ap.add_argument("--no-parity", action="store_true", help="skip the 2-slice parity block (oracle fp32 + fp64 on the host)")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: blocking flat all-reduce instead of the overlapped buckets")
    ap.add_argument("--sync-bn", action="store_true", help="N > 1: global-batch BatchNorm statistics (parallel.attach(sync_bn=True))")
    ap.add_argument("--graph", action="store_true",
                    help="N = 1: capture set_input + update() into ONE CUDA graph and time replays (small batches are launch-bound)")
    ap.add_argument("--cascades", type=int, default=12)
    ap.add_argument("--reg", default="Rec", choices=["Rec", "Mixed"],
                    help="training mode of the step: Rec = the headline cfg2 step; Mixed = BASELINE cfg5 (adds NetG twice, "
                         "NetD, the GAN losses and the discriminator step, reference model.py:217-239)")
    ap.add_argument("--lncc-weight", type=float, default=0.0, help="cfg3: add lncc_loss(full, warped) * weight to the step")
    ap.add_argument("--mi-weight", type=float, default=0.0, help="cfg5: add ms_mi_loss(full, warped) * weight to the step")
    ap.add_argument("--mask", default="equispaced", choices=["equispaced", "standard"])
    ap.add_argument("--sparsity", type=float, default=0.25)
    ap.add_argument("--checkpoint", default="auto", choices=["auto", "0", "1"],
                    help="recompute each cascade in backward (memory knob)")
    ap.add_argument("--cpu-sample", type=int, default=2, help="slices in the CPU baseline sample")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu = the contract's CPU arm; cuda = the same oracle port run as eager "
                         "PyTorch on the GPU (cuDNN convs, cuFFT; SURVEY 8d 'the reference on the B200 itself'), see --tf32")
    ap.add_argument("--tf32", type=int, default=0, help="--ref-device cuda: allow TF32 convs (the reference's own default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the instrumented per-kernel step")
    ap.add_argument("--breakdown", default="", help="write the per-op time table (JSON) to this path")
    ap.add_argument("--min-warmup", type=int, default=3, help="lower bound on warm-up steps (profiler runs only use < 3)")
    a = ap.parse_args()
    hw = str(a.shape).lower().split("x")
    a.H, a.W = (int(hw[0]), int(hw[-1]))
    a.shape = a.W                     # the column mask / num_low_frequencies live on the last axis (model.py:160-163)
    return a


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sust=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------- inputs
def make_inputs(batch, args, device="cpu", pin=False):
    """"rand" set of SURVEY.md 8(d): real & imag ~ U[0,1), complex64 (mirrors reference model.py:372-375)."""
    import torch
    g = torch.Generator().manual_seed(SEED + 1)
    shp = (batch, args.coils, args.H, args.W)
    full = torch.complex(torch.rand(*shp, generator=g), torch.rand(*shp, generator=g))
    aux = torch.complex(torch.rand(*shp, generator=g), torch.rand(*shp, generator=g))
    if pin:
        full, aux = full.pin_memory(), aux.pin_memory()
    if device != "cpu":
        full, aux = full.to(device), aux.to(device)
    return full, aux


def build_model(args):
    import torch
    from spatialalignmentnetwork_b200 import model as M
    torch.manual_seed(SEED)
    random.seed(SEED)
    cfg = M.Config(sparsity=args.sparsity, lr=1e-4, shape=args.shape, coils=args.coils, reg=args.reg, mask=args.mask,
                   weight_smooth=1000.0, weight_gan=0.1, weight_gan_sim=1.0, weight_sim=1.0, use_amp=False,
                   num_cascades=args.cascades)
    if args.lncc_weight:
        cfg.weight_lncc = args.lncc_weight
    if args.mi_weight:
        cfg.weight_mi = args.mi_weight
    net = M.CSModel(cfg)
    with torch.no_grad():  # non-trivial displacement field (zero-init makes offset == 0), SURVEY 8(d)
        torch.nn.init.normal_(net.net_T.net[-1].weight, 0, 1e-2)
    return net


# ------------------------------------------------------------------------------------- per-op model
def algorithmic(name, a):
    """(flops, bytes) of one C-ABI call from its scalar arguments (DESIGN.md 'algorithmic work')."""
    if name == "conv2d_fwd":
        N, Cin, H, W, Cout, K = a[4:10]
        return 2.0 * N * Cout * H * W * Cin * K * K, 4.0 * N * H * W * (Cin + Cout)
    if name == "conv2d_wgrad":
        N, Cin, H, W, Cout, K = a[4:10]
        return 2.0 * N * Cout * H * W * Cin * K * K, 4.0 * N * H * W * (Cin + Cout)
    if name == "tc_conv":
        N, H, W, Cin, Cout, K = a[4:10]
        return 2.0 * N * Cout * H * W * Cin * K * K, 4.0 * N * H * W * (Cin + Cout)
    if name == "tc_wgrad":
        N, H, W, Cin, Cout, K = a[5:11]
        return 2.0 * N * Cout * H * W * Cin * K * K, 4.0 * N * H * W * (Cin + Cout)
    if name == "fft_expand_dc":
        N, C, H, W = a[8:12]
        P = N * H * W
        return 0.0, ((32.0 * C + 8) * P if a[2] is not None else (16.0 * C + 8) * P)
    if name == "fft_reduce":
        N, C, H, W = a[5:9]
        P = N * H * W
        return 0.0, (16.0 * C + 8) * P + (8.0 * C * P if a[3] is not None else 0.0)
    if name == "fft_rss":
        N, C, H, W = a[4:8]
        return 0.0, (8.0 * C + 4) * N * H * W + (8.0 * C * N * H * W if a[2] is not None else 0.0)
    if name == "fft2":
        B, H, W = a[7:10]
        return 0.0, 16.0 * B * H * W
    return 0.0, 0.0


# dram__bytes_read.sum + dram__bytes_write.sum from the committed `ncu --set full` captures, per launch, next to the
# ALGORITHMIC bytes of SURVEY 8(d): a conv reads its fp32 input once and writes its fp32 output once,
# 4*N*H*W*(Cin+Cout).  The conv's real traffic includes the operand-staging pass that feeds it (fp32 in, BF16 hi/lo
# out) and the staged operand read back by the GEMM, so both kernels are counted.
NCU_TRAFFIC = {
    "tc_conv": {"launch": "3x3 18->18 @320x320 bs64 (stage_simple_kernel + conv_tc_kernel with the statistics epilogue)",
                "dram_bytes": (472.0e6 + 585.2e6) + (646.2e6 + 438.9e6),
                "algorithmic_bytes": 4.0 * 64 * 320 * 320 * (18 + 18),
                "source": "profiles/r2v_ncu_summary_stage_conv_wgrad.txt"},
    "tc_wgrad": {"launch": "3x3 18->18 @320x320 bs64 (wgrad_tc_kernel on the two staged operands)",
                 "dram_bytes": 1270.2e6 + 4.3e6,
                 "algorithmic_bytes": 4.0 * 64 * 320 * 320 * (18 + 18),
                 "source": "profiles/r2v_ncu_summary_stage_conv_wgrad.txt"},
    # fft_rows_v2_kernel + fft_cols_tma_kernel of one soft-DC fft_expand_dc call, bs 64, 320x320, 1 coil: (32C+8)*P algorithmic
    "fft_expand_dc": {"launch": "fft_expand_dc (soft DC) bs64 320x320 C=1 (fft_rows_v2_kernel + fft_cols_tma_kernel)",
                      "dram_bytes": (104.9e6 + 17.6e6) + (157.3e6 + 31.5e6), "algorithmic_bytes": 40.0 * 64 * 320 * 320,
                      "source": "profiles/r2r_fft_ncu_summary.txt"},
}

CLASSES = {
    "conv": ("conv2d_fwd", "conv2d_wgrad", "conv_pack_weights", "tc_conv", "tc_wgrad", "tc_stage_weights"),
    "operand_staging": ("tc_stage_act", "tc_stage_terms", "tc_unstage_act", "absmax"),
    "fft_dc": ("fft_expand_dc", "fft_reduce", "fft_rss", "fft2", "dc_bwd", "cmul_conj_planar"),
    "norm_act": ("plane_stats", "plane_stats_in", "in_stats_from_sums", "in_finalize_fwd", "bn_finalize_fwd", "affine_act_fwd", "act_bwd_reduce",
                 "in_finalize_bwd", "bn_finalize_bwd", "act_bwd_apply", "act_bwd_reduce_map", "act_bwd_apply_map",
                 "in_bwd_fused_map"),
    "resample": ("pool2", "up2", "depth_to_space2", "space_to_depth2", "axpby"),
    "align_warp": ("grid_from_offset", "grid_to_nchw", "warp_fwd", "warp_bwd", "grad_loss_fwd", "grad_loss_bwd"),
    "losses": ("ssim_loss_fwd", "ssim_loss_bwd", "lncc_loss_fwd", "lncc_loss_bwd", "mi_hist_fwd", "mi_hist_bwd",
               "filter2d", "pair_loss_fwd", "pair_loss_bwd"),
    "gan_weights": ("sn_sigma", "sn_scale", "sn_bwd"),
    "optimizer": ("adamw_step",),
    "metrics": ("error_sums", "mi_metric"),
}


def summarise_profile(records, step_ms, peaks):
    """records: [(name, args, ms)] of one instrumented step -> (roofline dict, breakdown dict)."""
    per = {}
    dc_fwd = dict(launches=0, ms=0.0, bytes=0.0)     # fft_expand_dc launches WITH the soft-DC epilogue (forward of a cascade)
    for name, a, ms in records:
        if name == "tc_conv_stats":        # the same conv_tc_kernel, with the statistics epilogue
            name = "tc_conv"
        fl, by = algorithmic(name, a)
        if name == "fft_expand_dc" and a[2] is not None:
            dc_fwd["launches"] += 1; dc_fwd["ms"] += ms; dc_fwd["bytes"] += by
        d = per.setdefault(name, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        d["launches"] += 1; d["ms"] += ms; d["flops"] += fl; d["bytes"] += by
    total = sum(d["ms"] for d in per.values())
    classes = {}
    for cname, members in CLASSES.items():
        ms = sum(per[m]["ms"] for m in members if m in per)
        classes[cname] = dict(ms=round(ms, 3), share_of_kernel_time=round(ms / total, 4) if total else 0.0)
    breakdown = dict(step_ms_instrumented=round(step_ms, 3), kernel_ms_total=round(total, 3), classes=classes,
                     ops={k: dict(launches=v["launches"], ms=round(v["ms"], 3),
                                  tflops=round(v["flops"] / v["ms"] / 1e9, 2) if v["flops"] and v["ms"] else None,
                                  gbs=round(v["bytes"] / v["ms"] / 1e6, 1) if v["bytes"] and v["ms"] else None)
                          for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"])})
    # dominant kernel = the U-Net convolution kernel with the largest share of the step
    roof = None
    labels = {"tc_conv": "conv_tc_kernel (tcgen05 implicit-GEMM conv on fp16-pair operands, 3 MMAs per product; fwd + dgrad)",
              "tc_wgrad": "wgrad_tc_kernel (tcgen05 weight gradient on fp16-pair operands)",
              "conv2d_fwd": "conv_fwd_kernel (fp32 CUDA-core conv, fwd + dgrad)",
              "conv2d_wgrad": "conv_wgrad_kernel (fp32 CUDA-core weight gradient)"}
    cands = [(per[k]["ms"], k) for k in labels if k in per and per[k]["ms"] > 0]
    if cands:
        _, key = max(cands)
        c = per[key]
        ach = c["flops"] / c["ms"] / 1e9  # TFLOP/s, algorithmic (1x) FLOPs: the BF16x3 split issues 3x this
        roof = dict(kernel=labels[key], bound="tensor", achieved=round(ach, 2),
                    peak=peaks["tf_sust"], unit="TFLOP/s", frac=round(ach / peaks["tf_sust"], 5),
                    peak_source=f"bf16 dense sustained, {peaks['src']}",
                    traffic=NCU_TRAFFIC.get(key),
                    launches=c["launches"], avg_launch_ms=round(c["ms"] / c["launches"], 4),
                    share_of_kernel_time=round(c["ms"] / total, 4),
                    issued_mma_frac=round(3 * ach / peaks["tf_sust"], 5) if key.startswith("tc_") else None,
                    hbm_gbs_compulsory=round(c["bytes"] / c["ms"] / 1e6, 1))
    f = per.get("fft_expand_dc")
    roof_fft = None
    if f and f["ms"] > 0:
        ach = f["bytes"] / f["ms"] / 1e6
        roof_fft = dict(kernel="fft_expand_dc (x*S -> fft2 -> soft-DC + residual; and its adjoint use)", bound="hbm",
                        achieved=round(ach, 1), peak=peaks["hbm"], unit="GB/s", frac=round(ach / peaks["hbm"], 4),
                        peak_source=f"copy bandwidth, {peaks['src']}", traffic=NCU_TRAFFIC.get("fft_expand_dc"),
                        launches=f["launches"],
                        avg_launch_ms=round(f["ms"] / f["launches"], 4),
                        share_of_kernel_time=round(f["ms"] / total, 4))
        if dc_fwd["launches"] and f["launches"] > dc_fwd["launches"]:
            # `achieved` above mixes the two uses of the entry point; split: the soft-DC launches (40 B per pixel and coil) and
            # the adjoint / expand-only launches of the backward (24 B: plain complex store, no k / k0 reads)
            adj_ms, adj_by, adj_n = f["ms"] - dc_fwd["ms"], f["bytes"] - dc_fwd["bytes"], f["launches"] - dc_fwd["launches"]
            roof_fft["soft_dc_launches"] = dict(launches=dc_fwd["launches"], avg_launch_ms=round(dc_fwd["ms"] / dc_fwd["launches"], 4),
                                                achieved=round(dc_fwd["bytes"] / dc_fwd["ms"] / 1e6, 1),
                                                frac=round(dc_fwd["bytes"] / dc_fwd["ms"] / 1e6 / peaks["hbm"], 4))
            roof_fft["expand_only_launches"] = dict(launches=adj_n, avg_launch_ms=round(adj_ms / adj_n, 4),
                                                    achieved=round(adj_by / adj_ms / 1e6, 1),
                                                    frac=round(adj_by / adj_ms / 1e6 / peaks["hbm"], 4))
    return roof, roof_fft, breakdown


# ------------------------------------------------------------------------------------- CPU oracle arm
def cpu_step_factory(args, nslices, device="cpu"):
    """The reference's path restated on CPU (oracle/): set_input + forwardT + forwardR + backward +
    AdamW step, on ``nslices`` slices, all host threads.  (``device='cuda'``: the same eager PyTorch ops on
    the GPU, i.e. what the reference itself would run there.)"""
    import torch
    from oracle import gan as ogan, step as ostep
    torch.set_num_threads(os.cpu_count() or 1)
    net = build_model(args)
    pruned = net.net_mask.pruned.clone().to(device)
    sds, opts = {}, {}
    for t in "TRGD":
        sd = {k: v.detach().clone().to(device) for k, v in getattr(net, "net_" + t).state_dict().items()}
        params = []
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k and "weight_u" not in k and "weight_v" not in k:
                v.requires_grad_(True)
                params.append(v)
        sds[t], opts[t] = sd, torch.optim.AdamW(params, lr=1e-4, weight_decay=0)
    full, aux = make_inputs(nslices, args, device=device)

    def step_rec():
        inp = ostep.set_input(full, aux, pruned)
        out = ostep.rec_step(sds["T"], sds["R"], inp, pruned, args.shape, args.sparsity, args.cascades,
                             weight_lncc=args.lncc_weight, weight_mi=args.mi_weight)
        opts["T"].zero_grad(); opts["R"].zero_grad()
        out["loss_all"].backward()
        opts["T"].step(); opts["R"].step()
        return out["loss_all"].item()

    def step_mixed():           # reference model.py:217-239: T, G, R step, then the discriminator step
        inp = ostep.set_input(full, aux, pruned)
        out = ogan.mixed_step(sds["T"], sds["R"], sds["G"], sds["D"], inp, pruned, args.shape, args.sparsity, args.cascades,
                              g_levels=4, d_blocks=(2, 2, 2, 2, 2), weight_lncc=args.lncc_weight, weight_mi=args.mi_weight)
        for t in "TGRD":
            opts[t].zero_grad()
        out["loss_G"].backward()
        for t in "TGR":
            opts[t].step()
        d = out["d_side"]()
        opts["D"].zero_grad()
        d["loss_D"].backward()
        opts["D"].step()
        return d["loss_D"].item()

    return step_mixed if args.reg == "Mixed" else step_rec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns = args.cpu_sample
    dev = args.ref_device
    if dev == "cuda":
        import torch
        torch.backends.cudnn.allow_tf32 = bool(args.tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
        dev = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
    step = cpu_step_factory(args, ns, device=dev)
    nwarm = args.warmup             # the same W warm-up and K timed steps as the b200 arm (a step is ~1.5 s on 2 slices)
    for _ in range(nwarm):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()                      # ends with .item(): synchronises the device arm too
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = ns / (ms / 1e3)
    cores = os.cpu_count() or 1
    sample = (f"{ns} slices/step of the same workload ({args.cascades}-cascade VarNet + align, {args.H}x{args.W}"
              f"{' x %d coils' % args.coils if args.coils > 1 else ''}, fwd+bwd+AdamW), CPU oracle port (torch CPU fp32), "
              f"{cores} threads; {nwarm} warm-up step(s)")
    if args.ref_device == "cuda":
        sample = sample.replace("CPU oracle port (torch CPU fp32)",
                                f"oracle port as eager PyTorch CUDA (cuDNN / cuFFT, TF32 {'on' if args.tf32 else 'off'})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "slices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": nwarm, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1), checkpoint=False, sample=ns),
        "ranks_running": 1,   # under torchrun only rank 0 runs this arm: its value does NOT scale with --gpus
        "cpu_baseline": {"value": round(val, 4), "unit": "slices/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def per_gpu_batch(args, world):
    if args.scaling == "strong":
        assert args.batch % world == 0, "--scaling strong: --batch (global) must divide over the GPUs"
        return args.batch // world
    return args.batch


def workload_config(args, world, checkpoint, sample=None):
    """``sample``: slices per step the arm REALLY ran when that is a bounded sample of the workload (reference arm)."""
    bpg = per_gpu_batch(args, world)
    dims = f"{args.H}x{args.W}" + (f" x {args.coils} coils" if args.coils > 1 else "")
    if args.reg == "Mixed":
        what = (f"cfg5 step: bs={bpg}/GPU {dims} synthetic T1/T2 pairs, {args.cascades}-cascade "
                "VarNet + alignment U-Net + NetG (x2) + NetD, reg='Mixed' (smooth*1000 + L1 gan_sim + SSIM + hinge*0.1, "
                "then the discriminator step), 4x equispaced mask, fwd+bwd+AdamW x4")
    else:
        name = "cfg4" if args.coils > 1 else ("cfg3" if args.lncc_weight else "cfg2")
        what = (f"{name}: bs={bpg}/GPU {dims} synthetic T1/T2 pairs, "
                f"{args.cascades}-cascade VarNet + alignment U-Net, reg='Rec' (smooth*1000 + SSIM), "
                "4x equispaced mask, fwd+bwd+AdamW")
    if args.lncc_weight:
        what += f" + lncc_loss(full, warped)*{args.lncc_weight:g}"
    if args.mi_weight:
        what += f" + ms_mi_loss(full, warped)*{args.mi_weight:g}"
    if args.mask != "equispaced" or args.sparsity != 0.25:
        what = what.replace("4x equispaced mask", f"{1 / args.sparsity:g}x {args.mask} mask")
    step_mb = 2 * bpg * args.coils * args.H * args.W * 8 / 1e6
    cfg = {"workload": what, "reg": args.reg,
           "batch_per_gpu": bpg, "global_batch": bpg * world,
           "shape": args.shape if args.H == args.W else [args.H, args.W], "cascades": args.cascades, "coils": args.coils,
           "parallelism": f"dp{world}", "checkpoint_cascades": bool(checkpoint), "cuda_graph": bool(getattr(args, "graph", False)),
           "l2": f"inputs ({step_mb:.0f} MB/step) + activations (GBs) exceed the 126 MB L2; no flush needed"}
    if sample is not None:
        # the reference arm times a BOUNDED SAMPLE of the workload: say what really ran
        cfg.update({"batch_per_gpu": sample, "global_batch": sample, "workload_batch_per_gpu": bpg,
                    "sample_mismatch": sample != bpg,
                    "workload": what + f" [this arm ran a bounded sample: {sample} slices/step on ONE process]"})
    return cfg


# ------------------------------------------------------------------------------------- CUDA arm
def gpu_reference_recorded(args):
    """The reference's own path as eager PyTorch CUDA (cuDNN / cuFFT) on a B200 of this pool, measured with
    ``bench.py --impl reference --ref-device cuda`` (SURVEY 8d "the real bar"); recorded, not re-timed here: it needs
    the whole GPU (bs 64 does not fit 180 GB in eager autograd; bs 32 is the largest that does)."""
    if args.reg != "Rec" or args.coils != 1 or args.H != 320 or args.W != 320 or args.cascades != 12:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r2a_gpu_reference_eager.json")) as f:
            r = json.load(f)["results"]
        return {"unit": "slices/s", "batch": 32, "tf32_off": r["tf0_bs32"]["slices_per_s"], "tf32_on": r["tf1_bs32"]["slices_per_s"],
                "bs64": "OOM (176 GiB)", "bs4_tf32_off": r["tf0_bs4"]["slices_per_s"],
                "source": "profiles/r2a_gpu_reference_eager.json (recorded run on a B200 of this pool, 5 steps, CUDA-synchronised wall clock)"}
    except Exception:
        return None


def parity_block(args, net, dev, nslices=2):
    """Headline-config parity inside the bench run: the same network (current weights) steps once on ``nslices``
    slices; outputs and EVERY parameter gradient are compared with the CPU oracle in fp32 and fp64, plus PSNR / SSIM
    of img_rec against the oracle's (BASELINE.json metric).  oracle/parity.py is test infrastructure: checker only."""
    import torch
    from oracle import parity
    full, aux = make_inputs(nslices, args)
    sd_T = {k: v.detach().cpu().clone() for k, v in net.net_T.state_dict().items()}
    sd_R = {k: v.detach().cpu().clone() for k, v in net.net_R.state_dict().items()}
    pruned = net.net_mask.pruned.detach().cpu().clone()
    net.train()
    net.set_input(full.to(dev), aux.to(dev))
    net.loss_all = 0
    net.forwardT()
    net.forwardR()
    for p in list(net.net_T.parameters()) + list(net.net_R.parameters()):
        p.grad = None
    net.loss_all.backward()
    torch.cuda.synchronize()
    out = {k: getattr(net, k).detach().cpu() for k in ("img_rec", "img_warped", "img_offset")}
    out["loss_all"] = net.loss_all.item()
    grads = {"T." + k: p.grad.detach().cpu() for k, p in net.net_T.named_parameters() if p.grad is not None}
    grads.update({"R." + k: p.grad.detach().cpu() for k, p in net.net_R.named_parameters() if p.grad is not None})
    rep = parity.rec_step_report(sd_T, sd_R, full, aux, pruned, args.shape, args.sparsity, args.cascades, out, grads)
    fw, g = rep["forward"], rep["grad"]
    return {"config": rep["config"], "oracle": "CPU restatement of the reference (oracle/), fp32 = reference arithmetic, fp64 = calibration",
            "forward_rel_l2_vs_fp32": {k: float(f"{v['vs_fp32']:.3e}") for k, v in fw.items()},
            "forward_fp32_oracle_vs_fp64": {k: float(f"{v['fp32_vs_fp64']:.3e}") for k, v in fw.items()},
            "grad_all_params_vs_fp64": float(f"{g['all_concatenated']['vs_fp64']:.3e}"),
            "grad_cpu_fp32_oracle_vs_fp64": float(f"{g['all_concatenated']['fp32_vs_fp64']:.3e}"),
            "grad_cosine_vs_fp64": round(g["all_concatenated"]["cosine_vs_fp64"], 6),
            "grad_cpu_fp32_oracle_cosine_vs_fp64": round(g["all_concatenated"]["fp32_cosine_vs_fp64"], 6),
            "grad_tensors": g["tensors"], "grad_worst": g["worst"][:3],
            "psnr_rec_vs_reference_rec_db": round(rep["image_metrics"]["psnr_rec_vs_reference_rec_db"], 2),
            "ssim_rec_vs_reference_rec": round(rep["image_metrics"]["ssim_rec_vs_reference_rec"], 7),
            "psnr_vs_full": {"b200": round(rep["image_metrics"]["psnr_rec_vs_full"], 3),
                             "reference": round(rep["image_metrics"]["psnr_reference_rec_vs_full"], 3)},
            "ssim_vs_full": {"b200": round(rep["image_metrics"]["ssim_rec_vs_full"], 5),
                             "reference": round(rep["image_metrics"]["ssim_reference_rec_vs_full"], 5)},
            "oracle_seconds": rep["oracle_seconds"]}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from spatialalignmentnetwork_b200 import _lib, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the san_b200 path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    net = build_model(args)
    total_mem = torch.cuda.get_device_properties(dev).total_memory
    bpg = per_gpu_batch(args, world)
    est = bpg * (args.cascades * 0.08 + 0.45) * 2 ** 30 * (args.H * args.W / 102400.0) * (0.5 + 0.5 * args.coils)   # raw conv outputs only
    ckpt = {"auto": est > 0.7 * total_mem, "0": False, "1": True}[args.checkpoint]
    net.net_R.checkpoint_cascades = ckpt
    net.to(dev).train()
    if world > 1:
        parallel.attach(net, overlap=not args.no_overlap, sync_bn=args.sync_bn)

    # per-rank shard of the global batch (weak: args.batch slices per GPU; strong: args.batch / world)
    full_h, aux_h = make_inputs(bpg, args, pin=True)
    if rank:
        full_h, aux_h = full_h.roll(rank, 0).pin_memory(), aux_h.roll(rank, 0).pin_memory()
    full_d, aux_d = full_h.to(dev), aux_h.to(dev)

    def step_resident():
        net.set_input(full_d, aux_d)
        net.update()

    # end-to-end step: every step moves one batch from pinned host memory (H2D inside the timed region) through the
    # package's input pipeline (prefetch.HostPrefetcher = the reference's pin_memory + non_blocking copy, train.py:150-165,
    # 207: the copy of batch k+1 runs on a copy stream while step k computes) and reads the step's loss back
    from spatialalignmentnetwork_b200.prefetch import HostPrefetcher

    def _host_batches():
        while True:
            yield (full_h, aux_h)
    feeder = [None]

    def step_e2e():
        if feeder[0] is None:
            feeder[0] = HostPrefetcher(_host_batches(), dev, depth=2)
        f, a = next(feeder[0])
        net.set_input(f, a)
        net.update()
        return net.loss_sim.item()      # D2H of the step's result (update() releases loss_all like the reference)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, args.min_warmup)):
        step_resident()
    gstep = None
    if args.graph:
        assert world == 1, "--graph: single GPU (the gradient all-reduce is launched from autograd hooks, not captured)"
        from spatialalignmentnetwork_b200.graphs import GraphedUpdate
        gstep = GraphedUpdate(net, full_d, aux_d, warmup=1)

        def step_resident():                    # noqa: F811  (inputs already in the graph's static buffers)
            gstep.replay()

        def step_e2e():                         # noqa: F811
            gstep(full_h, aux_h)                # H2D into the static buffers + replay
            return gstep.loss_sim.item()
        for _ in range(2):
            step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count()
    ms_res = timed(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    if gstep is not None:
        launches = gstep.launches_per_step * args.steps          # replays launch the captured kernels without host calls
    if gstep is None:
        feeder[0] = HostPrefetcher(_host_batches(), dev, depth=2)    # batch 0 in flight; every timed step issues one more copy
    for _ in range(2):          # untimed: the copy stream's device buffers come from cudaMalloc the first time round
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None
    peak_mem = torch.cuda.max_memory_allocated(dev)

    roof = roof_fft = breakdown = None
    if not args.no_profile:
        if gstep is not None:                   # the instrumented step runs eagerly (per-launch CUDA events)
            def step_resident():                # noqa: F811
                net.set_input(full_d, aux_d)
                net.update()
        barrier()
        # per-launch durations are only meaningful when kernels do not overlap: the instrumented step keeps the weight
        # gradient on the main stream (the timed steps above ran it on the side stream, next to the element-wise backward)
        from spatialalignmentnetwork_b200 import tc as _tc
        wg_overlap, _tc._WG_OVERLAP = _tc._WG_OVERLAP, False
        _lib.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident()
        e1.record()
        torch.cuda.synchronize()
        recs = _lib.profile_end()
        _tc._WG_OVERLAP = wg_overlap
        if rank == 0:
            roof, roof_fft, breakdown = summarise_profile(recs, e0.elapsed_time(e1), peaks)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = args.cpu_sample
        cstep = cpu_step_factory(args, ns)
        t0 = time.perf_counter()
        cstep()
        dt = time.perf_counter() - t0
        cores = os.cpu_count() or 1
        cpu_base = {"value": round(ns / dt, 4), "unit": "slices/s", "cores": cores, "kind": "port",
                    "sample": f"1 step on {ns} slices of the same workload (CPU oracle port, torch CPU fp32, {cores} threads, "
                              f"{dt:.1f} s)"}

    par = None
    if rank == 0 and world == 1 and not args.no_parity and args.reg == "Rec" and not args.lncc_weight and not args.mi_weight:
        par = parity_block(args, net, dev)

    if rank == 0:
        gb = bpg * world
        val = gb * args.steps / (ms_res / 1e3)
        e2e = gb * args.steps / (ms_e2e / 1e3)
        inbytes = (full_h.numel() * 8 + aux_h.numel() * 8) * world
        out = {"metric": METRIC, "value": round(val, 3), "unit": "slices/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, args.min_warmup), "ms_per_step": round(ms_res / args.steps, 3), "higher_is_better": True,
               "scaling": args.scaling, "vs_baseline": None,
               "dtype": "f16x3 (convs: fp16 hi/lo pair operands, 3 tcgen05 kind::f16 MMAs per product, fp32 TMEM accumulate = fp32-class products; fp32 storage, norms and FFT)",
               "data": "synthetic", "config": workload_config(args, world, ckpt),
               "e2e": {"value": round(e2e, 3), "unit": "slices/s", "h2d_bytes_per_step": inbytes, "d2h_bytes_per_step": 4,
                       "ms_per_step": round(ms_e2e / args.steps, 3),
                       "input_pipeline": ("H2D into the graph's static buffers, then replay" if gstep is not None else
                                          "prefetch.HostPrefetcher: one pinned-host -> device copy per step on a copy stream, "
                                          "overlapping the previous step; loss_sim.item() every step")},
               "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_fft_dc": roof_fft,
               "cpu_baseline": cpu_base, "gpu_reference": gpu_reference_recorded(args), "parity": par,
               "peak_mem_gb": round(peak_mem / 2 ** 30, 2), "impl": "b200"}
        if world > 1:
            gbk = getattr(net, "grad_buckets", None)
            out["grad_allreduce"] = ({"mode": "bucketed by cascade, launched from autograd hooks during backward",
                                      "buckets_launched_in_backward": gbk.launched} if gbk is not None
                                     else {"mode": "one blocking flat all-reduce after backward"})
        if breakdown is not None:
            out["kernel_time_shares"] = {k: v["share_of_kernel_time"] for k, v in breakdown["classes"].items()}
            if args.breakdown:
                os.makedirs(os.path.dirname(os.path.abspath(args.breakdown)), exist_ok=True)
                with open(args.breakdown, "w") as f:
                    json.dump(breakdown, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
