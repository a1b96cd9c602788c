#!/usr/bin/env python
"""BASELINE configs[3]: Pitts250k-shaped run -- a synthetic gallery sharded across the GPUs of one box.

    python tools/bench_gallery.py --n-db 250000 --n-q 6800                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_gallery.py ...

Every rank extracts its contiguous slice of the database (DistributedSliceSampler semantics: ceil(n/W) images,
wrap-around padding masked out later) and of the queries with the full VGG16 + NetVLAD + PCA(4096) path; the
descriptor shard stays in that GPU's HBM (n/W x 16 KiB).  Queries are all-gathered (6.8k x 16 KiB = 111 MB),
every rank ranks all queries against its shard (tcgen05 distance + top-10 + exact re-scoring), and ONE NCCL
all-gather of the [m,10] candidates + a merge kernel gives the final ranking on every rank.

Images are generated on the device (torch.randn per batch, seeded per rank) -- 250k x 3.7 MB of host images is
not a workload this box can hold; the host->device leg is measured by bench.py's `e2e` instead.  Strong scaling:
the same total work on 1..8 GPUs; time = max over ranks of CUDA-event time.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openibl_b200 import synth  # noqa: E402
from openibl_b200.engine import Engine  # noqa: E402
from openibl_b200.evaluators import sharded_topk  # noqa: E402
from openibl_b200.utils.data.sampler import slice_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-db", type=int, default=250000)
    ap.add_argument("--n-q", type=int, default=6800)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
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

    def extract_slice(n_total, seed):
        lo, cnt, per = slice_bounds(n_total, world, rank)
        out = torch.zeros(per, 4096, device=dev)
        g = torch.Generator(device=dev).manual_seed(seed * 1000 + rank)
        for b0 in range(0, cnt, args.batch):
            nb = min(args.batch, cnt - b0)
            x = torch.randn(nb, 3, args.height, args.width, device=dev, generator=g)
            d, _ = eng.extract(x, pca=True)
            out[b0:b0 + nb] = d
        return out, lo, cnt, per

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (allocations, TMA descriptors, clocks)
    eng.extract(torch.randn(args.batch, 3, args.height, args.width, device=dev), pca=True)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    shard, lo, cnt, per = extract_slice(args.n_db, seed=1)
    qslice, qlo, qcnt, qper = extract_slice(args.n_q, seed=2)
    ev[1].record()
    if world > 1:
        parts = [torch.empty_like(qslice) for _ in range(world)]
        dist.all_gather(parts, qslice)
        q = torch.cat(parts)[: args.n_q].contiguous()
    else:
        q = qslice[: args.n_q].contiguous()
    ev[2].record()
    dk, ik = sharded_topk(q, shard, 10, idx_base=lo, n_valid=cnt)
    ev[3].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]),
                      ev[0].elapsed_time(ev[3])], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    chk = float(dk.double().sum().item())
    ok = bool((ik >= 0).all() and (ik < args.n_db).all() and (dk[:, 1:] >= dk[:, :-1]).all())
    if rank == 0:
        ext_ms, gather_ms, rank_ms, total_ms = [float(v) for v in t.tolist()]
        print(json.dumps({
            "workload": f"{args.n_db} db + {args.n_q} query images {args.height}x{args.width}, VGG16+NetVLAD+PCA4096, "
                        f"top-10 ranking, {world} GPU(s), db shard {per} rows/GPU",
            "n_gpus": world, "extract_s": ext_ms / 1e3, "query_allgather_ms": gather_ms, "ranking_ms": rank_ms,
            "total_s": total_ms / 1e3, "images_per_s": (args.n_db + args.n_q) / (ext_ms / 1e3),
            "pairs_per_s": args.n_q * args.n_db / (rank_ms / 1e3), "topk_sane": ok, "topk_dist_checksum": chk,
            "launches": eng.launch_count}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
