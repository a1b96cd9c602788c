#!/usr/bin/env python
"""BASELINE configs[3]: Pitts250k-shaped run -- a synthetic gallery sharded across the GPUs of one box
(openibl_b200/gallery.py holds the flow; this is its command line).

    python tools/bench_gallery.py --n-db 250000 --n-q 6800                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_gallery.py ...
    python tools/bench_gallery.py --emulate-world 8 ...                         # all 8 ranks played on one GPU

The gallery does not depend on the world size (images are seeded by global index), so `topk_index_hash` and
`recalls` of a 1-GPU and an 8-GPU run must be EQUAL; `exact_fp64_subset_agreement` compares a query subset with
an fp64 ranking.  Strong scaling: the same total work on 1..8 GPUs; times are max over ranks of CUDA-event
times, NCCL warmed up before the timed region.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openibl_b200 import gallery, synth  # noqa: E402
from openibl_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-db", type=int, default=250000)
    ap.add_argument("--n-q", type=int, default=6800)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--emulate-world", type=int, default=0)
    ap.add_argument("--no-exact", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine.get(local)
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(seed=0, sharp=True, with_pca=True).items()}
    slots = synth.VGG16_CONV_SLOTS
    eng.set_vgg16([sd[f"base_model.base.{s}.weight"] for s in slots], [sd[f"base_model.base.{s}.bias"] for s in slots])
    eng.set_netvlad(sd["net_vlad.conv.weight"], sd["net_vlad.centroids"])
    eng.set_pca(sd["pca_layer.weight"], sd["pca_layer.bias"])
    gallery.center_pca(eng, sd["pca_layer.weight"], args.height, args.width, args.batch)   # bias = -W.mean, as a PCA fit sets it
    # warm-up outside the timed region: allocations, TMA descriptors, clocks, and the NCCL communicator
    gallery.run(eng, 4 * args.batch * max(world, args.emulate_world, 1), 64, args.height, args.width, args.batch,
                emulate_world=args.emulate_world, check_exact=False)
    res = gallery.run(eng, args.n_db, args.n_q, args.height, args.width, args.batch, emulate_world=args.emulate_world,
                      check_exact=not args.no_exact)
    if rank == 0:
        res["launches"] = eng.launch_count
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
