#!/usr/bin/env python
"""BASELINE configs[4]: ONE SFRS training step (reference examples/netvlad_img_sfrs.py + ibl/trainers.py:181-259) under
DistributedDataParallel, one process per GPU, on synthetic tuples -- exercises the NetVLAD backward, the conv5 dgrad /
wgrad kernels and DDP's gradient all-reduce over NCCL.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/sfrs_step_synthetic.py \
        --launcher pytorch --tuple-size 4 --neg-num 10 --diff-num 10 --height 480 --width 640

Model setup as netvlad_img_sfrs.py:96-115 (vgg16 with train_layers='conv5' semantics: everything below conv5 frozen;
NetVLAD with _init_params-style parameters; EmbedRegionNet; DDP with find_unused_parameters=True), a frozen teacher
`model_cache`, SGD(lr, momentum, weight_decay) on the trainable parameters.  Rank r trains on its own tuples
(seed + r).  Prints one JSON line from rank 0: the losses of the step, its device time, and whether every rank holds
identical parameters after the optimizer step (the DDP invariant)."""
from __future__ import print_function, absolute_import

import argparse
import json
import os.path as osp
import sys

import torch
from torch import nn

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))

from ibl import models  # noqa: E402
from ibl.trainers import SFRSTrainer  # noqa: E402
from ibl.utils.dist_utils import init_dist, synchronize  # noqa: E402
from openibl_b200 import synth  # noqa: E402


def build(seed, args, ddp):
    sd = synth.make_state_dict(seed=seed, sharp=True, with_pca=False, bias_scale=0.02)
    base = models.create("vgg16", pretrained=False, train_layers="conv5")
    pool = models.create("netvlad", dim=base.feature_dim)
    model = models.create("embedregionnet", base, pool, tuple_size=args.tuple_size)
    model.load_state_dict(sd)
    for layer in list(model.base_model.base.children())[:24]:     # vgg.py:50-53 with train_layers='conv5'
        for p in layer.parameters():
            p.requires_grad = False
    model.cuda(args.gpu)
    if ddp:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[args.gpu], output_device=args.gpu,
                                                    find_unused_parameters=True)
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launcher", default="pytorch", choices=["pytorch", "slurm"])
    ap.add_argument("--tcp-port", default="5017")
    ap.add_argument("--tuple-size", type=int, default=4)
    ap.add_argument("--neg-num", type=int, default=10)
    ap.add_argument("--diff-num", type=int, default=10)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--generation", type=int, default=0)
    ap.add_argument("--seed", type=int, default=31)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--lr", type=float, default=0.001)
    args = ap.parse_args()
    init_dist(args.launcher, args)
    synchronize()
    world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
    model = build(13, args, ddp=True)
    cache = build(23, args, ddp=False)
    trainer = SFRSTrainer(model, cache, margin=0.1, neg_num=args.neg_num, gpu=args.gpu, temp=[0.07, 0.07, 0.06, 0.05])
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=args.lr, momentum=0.9, weight_decay=0.001)
    easy, diff = synth.make_sfrs_tuples(seed=args.seed + rank, tuples=args.tuple_size, neg_num=args.neg_num,
                                        n_diff=args.diff_num, height=args.height, width=args.width)
    easy, diff = easy.cuda(args.gpu), diff.cuda(args.gpu)
    model.train()
    cache.train()
    out = {}
    for step in range(args.steps + 1):                       # step 0 = warm-up (allocations, NCCL), not applied
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        synchronize()
        e0.record()
        loss_hard, loss_soft = trainer._forward(easy, diff, "sare_ind", args.generation)
        loss = loss_hard + 0.5 * loss_soft
        opt.zero_grad()
        loss.backward()
        if step > 0:
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        out = {"loss_hard": float(loss_hard), "loss_soft": float(loss_soft), "step_ms": e0.elapsed_time(e1)}
    gnorm = float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in params if p.grad is not None)))
    chk = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    torch.distributed.all_gather(allc, chk)
    t = torch.tensor([out["step_ms"]], device=chk.device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        n_img = args.tuple_size * (2 * (1 + args.diff_num) + 2 + args.neg_num)
        print("SFRS_STEP " + json.dumps({
            "world": world, "tuple_size": args.tuple_size, "neg_num": args.neg_num, "diff_num": args.diff_num,
            "image": [args.height, args.width], "generation": args.generation, "loss_hard": out["loss_hard"],
            "loss_soft": out["loss_soft"], "step_ms_max_over_ranks": float(t.item()),
            "images_forward_per_gpu": n_img, "grad_norm_rank0": gnorm, "finite": bool(torch.isfinite(chk).all()),
            "params_identical_across_ranks": bool(all(torch.equal(c, allc[0]) for c in allc)),
            "trainable_params": int(sum(p.numel() for p in params))}), flush=True)
    synchronize()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
