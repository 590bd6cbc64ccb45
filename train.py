#!/usr/bin/env python
"""Training entry point with the reference's CLI (reference train.py:341-417) for the san_b200 path.

Same flags, same loop order (batch -> device -> set_input -> update -> periodic scalars / checkpoints;
reference train.py:198-253), same checkpoint directory format (basemodel.ckpt_save).  Differences:
  * one process per GPU under ``torchrun`` (the reference is single-GPU): the batch is sharded across ranks and
    gradients are all-reduced (spatialalignmentnetwork_b200.parallel);
  * ``--train synthetic[:N]`` / ``--val synthetic[:N]`` generate seeded phantom slice pairs on the device instead
    of reading the fastMRI h5 volumes (h5py and the data are not available offline); csv paths are accepted only
    when ``h5py`` is importable;
  * ``--use_amp`` is refused (fp32 parity path), ``--prefetch`` / ``--protocals`` are accepted and ignored for synthetic
    data, ``--intel_stop`` keeps ``<logdir>/best.pt`` like the reference;
  * ``--gan_layers_G`` / ``--gan_layers_D`` (not in the reference) shrink the GAN
    networks for smoke runs.
"""
import argparse
import json
import os
import random
import shutil
import time

import torch


def synthetic_pairs(n, shape, coils, device, seed):
    """Seeded T2-like phantom (sum of soft ellipses) and a T1-like contrast of it, complex64 with imag = 0
    (reference paired_dataset.py:69-73 casts magnitude images the same way)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, shape), torch.linspace(-1, 1, shape), indexing="ij")
    out = torch.zeros(n, 1, shape, shape)
    for i in range(n):
        for _ in range(12):
            cx, cy = (torch.rand(2, generator=g) - 0.5).tolist()
            ax, ay = (0.1 + 0.4 * torch.rand(2, generator=g)).tolist()
            amp = torch.rand(1, generator=g).item()
            out[i, 0] += amp * torch.sigmoid(8 * (1 - ((xx - cx) / ax) ** 2 - ((yy - cy) / ay) ** 2))
        out[i] /= out[i].max().clamp_min(1e-6)
    t2 = out.expand(n, coils, shape, shape).contiguous()
    t1 = (1 - 0.7 * t2 + 0.3 * t2 ** 2) * (t2 > 0.05)
    return torch.complex(t2, torch.zeros_like(t2)).to(device), torch.complex(t1, torch.zeros_like(t1)).to(device)


def main(args):
    assert not args.use_amp, "--use_amp: the san_b200 path computes in fp32 (BF16x3 tensor-core convs); no AMP mode"
    import torch.distributed as dist
    from spatialalignmentnetwork_b200 import parallel
    from spatialalignmentnetwork_b200.augment import augment_funcs, center_crop
    from spatialalignmentnetwork_b200.model import CSModel, Config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    random.seed(19950102 + 666 + 233)                       # reference train.py:99-100 (mask offset)
    torch.manual_seed(19950102)
    cfg = Config(sparsity=args.sparsity, lr=args.lr, shape=args.crop, coils=args.coils, reg=args.reg, mask=args.mask,
                 weight_smooth=args.smooth_weight, weight_gan=args.gan_weight, weight_gan_sim=args.gan_sim_weight,
                 weight_sim=args.sim_weight, use_amp=False, num_cascades=args.num_cascades)
    if args.lncc_weight:
        cfg.weight_lncc = args.lncc_weight
    if args.mi_weight:
        cfg.weight_mi = args.mi_weight
    if args.gan_layers_G:
        cfg.gan_layers_G = [int(c) for c in args.gan_layers_G.split(",")]
    if args.gan_layers_D:
        cfg.gan_layers_D = [[int(c) for c in blk.split(",")] for blk in args.gan_layers_D.split(";")]
    if args.resume:
        # the CLI config wins over the checkpoint's (reference train.py:121): the staged recipe of
        # commands_train_test.sh resumes each stage with a different --reg and --load_nets
        net = CSModel(ckpt=args.resume, cfg=cfg, objects=args.load_nets)
    else:
        assert args.load_nets is None, "--load_nets needs --resume (reference train.py:123)"
        net = CSModel(cfg)
    net.to(device)
    if world > 1:
        parallel.attach(net, sync_bn=args.sync_bn)
    assert args.train.startswith("synthetic") and args.val.startswith("synthetic"), \
        "offline build: use --train synthetic[:N] --val synthetic[:N] (h5 volumes are not available)"
    n_train = int(args.train.split(":")[1]) if ":" in args.train else 64
    n_val = int(args.val.split(":")[1]) if ":" in args.val else 16
    full_t, aux_t = synthetic_pairs(n_train, args.crop, args.coils, device, 1)
    full_v, aux_v = synthetic_pairs(n_val, args.crop, args.coils, device, 2)
    if rank == 0:
        os.makedirs(args.logdir, exist_ok=True)
    it, t0 = 0, time.time()
    loss_best, iter_best = None, 0
    for epoch in range(args.epoch):
        net.train()
        perm = torch.randperm(n_train, generator=torch.Generator().manual_seed(epoch)).to(device)
        for b0 in range(0, n_train - args.batch_size + 1, args.batch_size):
            idx = parallel.shard(perm[b0:b0 + args.batch_size], rank, world)
            with torch.no_grad():                           # reference train.py:208-210
                batch = augment_funcs[args.aux_aug]([full_t[idx], aux_t[idx]])
                batch = [center_crop(x, (net.cfg.shape, net.cfg.shape)).contiguous() for x in batch]
            net.set_input(*batch)
            net.update()
            it += 1
            if rank == 0 and it % args.log_every == 0:
                sc = net.get_vis("scalars")["scalars"]
                print(json.dumps({"iter": it, "epoch": epoch, "sec": round(time.time() - t0, 1),
                                  **{k: round(float(v), 6) for k, v in sc.items()}}), flush=True)
            if rank == 0 and it % args.ckpt_every == 0:
                net.save(os.path.join(args.logdir, f"ckpt_{it:010d}.pt"))
        # ---- validation (reference train.py:267-308): disjoint windows of the global batch size (drop_last like the
        # reference's val loader), each window sharded over the ranks; metric sums are all-reduced so that every rank
        # sees the same numbers and takes the same early-stop decision.  BatchNorm running statistics are rank-local
        # during training: they are averaged first so that eval-mode outputs (and the saved checkpoint) agree.
        parallel.average_buffers(net)
        net.eval()
        sums = torch.zeros(4, dtype=torch.float64, device=device)       # loss, PSNR, SSIM, windows
        for b0 in range(0, n_val - args.batch_size + 1, args.batch_size):
            fv = parallel.shard(full_v[b0:b0 + args.batch_size], rank, world)
            av = parallel.shard(aux_v[b0:b0 + args.batch_size], rank, world)
            if fv.shape[0] < 2:                             # forwardG splits the batch in two halves (model.py:125-126)
                continue
            net.set_input(fv, av)
            sums += torch.tensor([net.test(), net.metric_PSNR, net.metric_SSIM, 1.0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(sums)
        nwin = max(sums[3].item(), 1.0)
        loss_current, val_psnr, val_ssim = (sums[:3] / nwin).tolist()
        if rank == 0:
            print(json.dumps({"epoch": epoch, "val_PSNR": val_psnr, "val_SSIM": val_ssim}), flush=True)
        if args.intel_stop > 0:                             # early stopping of reference train.py:293-307
            if loss_best is None or loss_current < loss_best:
                loss_best, iter_best = loss_current, it
                if rank == 0:
                    best = os.path.join(args.logdir, "best.pt")
                    if os.path.exists(best):
                        shutil.rmtree(best)
                    net.save(best)
            elif it >= args.intel_stop + iter_best:         # identical on every rank (all-reduced loss): no rank is left
                if rank == 0:                               # waiting in a gradient all-reduce
                    print("signal_end set due to intel_stop", flush=True)
                break
    if rank == 0:
        net.save(os.path.join(args.logdir, f"ckpt_{it:010d}_final.pt"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    p = argparse.ArgumentParser(description="CS with adaptive mask (san_b200 path)")
    p.add_argument("--logdir", type=str, required=True)
    p.add_argument("--resume", type=str, default=None)
    p.add_argument("--load_nets", type=str, nargs="*", default=None)
    p.add_argument("--epoch", type=int, default=150)
    p.add_argument("--batch_size", type=int, default=10, help="GLOBAL batch size (sharded over ranks)")
    p.add_argument("--num_workers", type=int, default=0)
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--reg", type=str, required=True, choices=["None", "Rec", "Mixed", "GAN-Only"])
    p.add_argument("--smooth_weight", type=float, required=True)
    p.add_argument("--gan_weight", type=float, default=0.0)
    p.add_argument("--gan_sim_weight", type=float, default=0.0)
    p.add_argument("--sim_weight", type=float, required=True)
    p.add_argument("--mask", type=str, required=True)
    p.add_argument("--sparsity", type=float, default=None)
    p.add_argument("--train", type=str, required=True)
    p.add_argument("--val", type=str, required=True)
    p.add_argument("--crop", type=int, default=320)
    p.add_argument("--coils", type=int, default=1)
    p.add_argument("--aux_aug", type=str, default="None", choices=["None", "Rigid", "BSpline", "PBSpline"])
    p.add_argument("--force_gpu", action="store_true")
    p.add_argument("--intel_stop", type=int, default=0, metavar="N",
                   help="stop when the validation loss has not improved for N iterations; keeps <logdir>/best.pt")
    p.add_argument("--protocals", type=str, default=None, nargs="*", help="input modalities (h5 datasets only; ignored for synthetic data)")
    p.add_argument("--prefetch", action="store_true", help="accepted for CLI compatibility: synthetic data is already device-resident")
    p.add_argument("--use_amp", action="store_true", help="not supported: the san_b200 path is the fp32 parity path")
    p.add_argument("--num_cascades", type=int, default=8)
    p.add_argument("--lncc_weight", type=float, default=0.0, help="extra registration term lncc_loss(full, warped) (BASELINE cfg3)")
    p.add_argument("--mi_weight", type=float, default=0.0, help="extra registration term ms_mi_loss(full, warped) (BASELINE cfg5)")
    p.add_argument("--gan_layers_G", type=str, default="", help="e.g. 8,16,16 (default: the reference's 64,128,256,512,512)")
    p.add_argument("--gan_layers_D", type=str, default="", help="e.g. '8,8;16,16' (default: the reference's widths)")
    p.add_argument("--sync_bn", action="store_true",
                   help="torchrun only: BatchNorm statistics of the GLOBAL batch (the sharded step then equals the "
                        "single-process reference at --batch_size); default: per-rank statistics")
    p.add_argument("--log_every", type=int, default=50)
    p.add_argument("--ckpt_every", type=int, default=1000)
    main(p.parse_args())
