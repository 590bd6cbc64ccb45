#!/usr/bin/env python
"""Evaluation entry point (reference eval.py:29-86): load a checkpoint directory, run ``CSModel.test()`` on
batches of slices, print the mean metrics as JSON.  ``--val synthetic[:N]`` uses the seeded phantom pairs of
train.py (the h5 volumes are not available offline); ``--aux_aug FACTOR`` misaligns the auxiliary modality like the
reference's robustness evaluation (eval.py:15-27,43-58), ``--metric FILE`` stores the per-batch scalars
(eval.py:80-82); NIfTI export is out of scope (nibabel absent)."""
import argparse
import json

import torch


def main(args):
    from spatialalignmentnetwork_b200.model import CSModel
    from train import synthetic_pairs
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    net = CSModel(ckpt=args.resume)
    net.use_amp = False                                      # reference eval.py:41
    net.to(device).eval()
    n = int(args.val.split(":")[1]) if ":" in args.val else 16
    full, aux = synthetic_pairs(n, net.cfg.shape, net.cfg.coils, device, 2)
    rows = []
    for b0 in range(0, n, args.batch_size):
        batch = (full[b0:b0 + args.batch_size], aux[b0:b0 + args.batch_size])
        if args.aux_aug > 0:
            from spatialalignmentnetwork_b200.augment import augment_aux
            with torch.no_grad():
                batch = augment_aux(batch, args.aux_aug)
        net.set_input(*batch)
        net.test()
        rows.append({k: getattr(net, k) for k in ("metric_PSNR", "metric_SSIM", "metric_MAE", "metric_MSE", "metric_MI")})
    mean = {k: sum(r[k] for r in rows) / len(rows) for k in rows[0]}
    print(json.dumps(mean))
    if args.metric:
        with open(args.metric, "w") as f:
            json.dump(rows, f)
    if args.save:
        with open(args.save, "w") as f:
            json.dump({"mean": mean, "batches": rows}, f)


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--resume", type=str, required=True, help="checkpoint directory")
    p.add_argument("--val", type=str, default="synthetic:16")
    p.add_argument("--batch_size", type=int, default=8)
    p.add_argument("--save", type=str, default=None)
    p.add_argument("--metric", type=str, default=None, help="JSON file for the per-batch metrics")
    p.add_argument("--aux_aug", type=float, default=-1, help="> 0: misalign the auxiliary modality by this factor")
    p.add_argument("--crop", type=int, default=None, help="accepted for CLI compatibility (the checkpoint's shape is used)")
    p.add_argument("--protocals", type=str, default=None, nargs="*", help="accepted for CLI compatibility (h5 datasets only)")
    main(p.parse_args())
