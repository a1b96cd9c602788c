#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

TEST INFRASTRUCTURE.  Runs only in the build container, where /root/reference
exists (it does not exist on the GPU box; nothing at test/bench time reads it).
The reference imports h5py at package import (ibl/pca.py:10) although the
forward path never touches it; h5py is not installed here, so an empty stub
module is injected before the import.

    python oracle/gen_golden.py            # rewrites tests/golden/

Inputs/weights come from openibl_b200.synth (seeded), loaded into the reference
modules through load_state_dict, so the fixtures depend only on the seeds and
on torch's CPU kernels (torch 2.11.0 here).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("IBL_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, REF)

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

from ibl import models as ref_models  # noqa: E402  (the reference's ibl)
from ibl import evaluators as ref_eval  # noqa: E402
from openibl_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_ref_model(sd, with_pca, pca_dim=4096):
    base = ref_models.create("vgg16", pretrained=False)
    pool = ref_models.create("netvlad", dim=base.feature_dim)
    if with_pca:
        model = ref_models.create("embednetpca", base, pool, dim=pca_dim)
    else:
        model = ref_models.create("embednet", base, pool)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


@torch.no_grad()
def case_hub_480x640():
    """BASELINE config 1: hubconf.vgg16_netvlad(pretrained=False)-shaped model,
    one 3x480x640 image (hubconf.py:4-10 -> netvlad.py:95-110)."""
    sd = synth.make_state_dict(seed=0, with_pca=True)
    model = build_ref_model(sd, with_pca=True)
    x = synth.make_images(seed=1, batch=1)
    desc = model(x)
    pool, feat = model.base_model(x)
    raw = model.net_vlad(feat)
    sd_np = {k: v for k, v in sd.items() if not k.startswith("pca_layer")}
    emb = build_ref_model(sd_np, with_pca=False)
    pool2, vlad = emb(x)
    assert torch.equal(pool, pool2)
    save("hub_480x640",
         desc=desc.numpy(), vlad=vlad.numpy(), pool=pool.numpy(),
         raw_vlad_sub=raw[:, ::4, ::8].numpy(),
         feat_sub=feat[:, ::8, ::3, ::4].numpy(),
         feat_sum=np.float64(feat.double().sum().item()),
         feat_abs_sum=np.float64(feat.double().abs().sum().item()))


@torch.no_grad()
def case_small_96x128():
    """Small image, batch 2, non-zero biases: every stage boundary kept in full."""
    sd = synth.make_state_dict(seed=5, with_pca=True, pca_dim=128, bias_scale=0.05)
    model = build_ref_model(sd, with_pca=True, pca_dim=128)
    x = synth.make_images(seed=6, batch=2, height=96, width=128)
    desc = model(x)
    pool, feat = model.base_model(x)
    raw = model.net_vlad(feat)
    emb = build_ref_model({k: v for k, v in sd.items() if not k.startswith("pca_layer")}, with_pca=False)
    _, vlad = emb(x)
    save("small_96x128", desc=desc.numpy(), vlad=vlad.numpy(), pool=pool.numpy(),
         raw_vlad=raw.numpy(), feat=feat.numpy())


@torch.no_grad()
def case_odd_70x90():
    """Odd sizes: floor-mode pooling (70x90 -> 35x45 -> 17x22 -> 8x11 -> 4x5)."""
    sd = synth.make_state_dict(seed=7, with_pca=False, bias_scale=0.05)
    emb = build_ref_model(sd, with_pca=False)
    x = synth.make_images(seed=8, batch=1, height=70, width=90)
    pool, vlad = emb(x)
    _, feat = emb.base_model(x)
    save("odd_70x90", vlad=vlad.numpy(), pool=pool.numpy(), feat=feat.numpy())


@torch.no_grad()
def case_netvlad_unit():
    """NetVLAD layer alone on a 30x40 map (S=1200), default and sharp (_init_params)
    parameters (netvlad.py:34-61)."""
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(2, 512, 30, 40, generator=g) * 3.0 + 0.5
    out = {}
    for tag, sharp in (("soft", False), ("sharp", True)):
        p = synth.make_netvlad_params(seed=3, sharp=sharp)
        layer = ref_models.create("netvlad", dim=512)
        if sharp:
            # drive the reference's own _init_params (netvlad.py:34-42)
            gg = synth._gen(3 + 1000)
            clsts = torch.randn(64, 512, generator=gg)
            clsts = clsts / clsts.norm(dim=1, keepdim=True)
            desc = torch.randn(5000, 512, generator=gg)
            desc = desc / desc.norm(dim=1, keepdim=True)
            layer.clsts = clsts.numpy().astype(np.float32)
            layer.traindescs = desc.numpy().astype(np.float32)
            layer._init_params()
            assert torch.allclose(layer.conv.weight.data, p["conv_weight"], rtol=1e-6, atol=1e-6)
            assert abs(layer.alpha - p["alpha"]) < 1e-3 * p["alpha"]
            out["alpha"] = np.float64(layer.alpha)
        else:
            layer.centroids.data.copy_(p["centroids"])
            layer.conv.weight.data.copy_(p["conv_weight"])
        raw = layer.eval()(feat)
        v = torch.nn.functional.normalize(raw, p=2, dim=2).view(2, -1)
        v = torch.nn.functional.normalize(v, p=2, dim=1)
        z = layer.conv(torch.nn.functional.normalize(feat, p=2, dim=1)).view(2, 64, -1)
        out[f"{tag}_raw"] = raw.numpy()
        out[f"{tag}_vlad"] = v.numpy()
        out[f"{tag}_maxprob"] = np.float64(torch.softmax(z, 1).max().item())
    save("netvlad_unit", **out)


@torch.no_grad()
def case_pca_unit():
    """EmbedNetPCA.pca_layer + L2 (netvlad.py:105-108) on unit vectors, P=64."""
    p = synth.make_pca_params(seed=9, in_dim=32768, out_dim=64)
    g = torch.Generator().manual_seed(12)
    v = torch.nn.functional.normalize(torch.randn(5, 32768, generator=g), dim=1)
    conv = torch.nn.Conv2d(32768, 64, 1)
    conv.weight.data.copy_(p["weight"]); conv.bias.data.copy_(p["bias"])
    y = conv(v.view(5, 32768, 1, 1)).view(5, -1)
    y = torch.nn.functional.normalize(y, p=2, dim=-1)
    save("pca_unit", out=y.numpy())


def case_retrieval():
    """pairwise_distance + evaluate_all (+nms) from the reference (evaluators.py:105-167),
    run under a 1-rank gloo group because both call dist.get_rank() unconditionally."""
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)
    q, db, gt = synth.make_gallery(n_db=1500, n_q=300, dim=512, sigma=0.28)
    feats = {}
    query = [("q%05d" % i, i, 0.0, 0.0) for i in range(q.size(0))]
    # place ids: every 3 consecutive db images share a pid (exercises spatial_nms)
    gallery = [("d%05d" % i, i // 3, 0.0, 0.0) for i in range(db.size(0))]
    for (f, _, _, _), row in zip(query, q):
        feats[f] = row
    for (f, _, _, _), row in zip(gallery, db):
        feats[f] = row
    distmat, xq, yg = ref_eval.pairwise_distance(feats, query, gallery)
    gt_list = [np.array([int(g)]) for g in gt]
    top10 = np.argsort(distmat.numpy(), axis=1)[:, :10]
    rec = ref_eval.evaluate_all(distmat.clone(), gt_list, gallery)
    rec_nms = ref_eval.evaluate_all(distmat.clone(), gt_list, gallery, nms=True)
    self_d, _, _ = ref_eval.pairwise_distance({k: feats[k] for k in list(feats)[:64]})
    save("retrieval", dist_sub=distmat.numpy()[:32], top10=top10,
         top10_dist=np.take_along_axis(distmat.numpy(), top10, 1),
         recalls=rec, recalls_nms=rec_nms, self_dist=self_d.numpy(), gt=gt.numpy())
    dist.destroy_process_group()


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1:] or ["hub", "small", "odd", "netvlad", "pca", "retrieval"]
    if "small" in which: case_small_96x128()
    if "odd" in which: case_odd_70x90()
    if "netvlad" in which: case_netvlad_unit()
    if "pca" in which: case_pca_unit()
    if "retrieval" in which: case_retrieval()
    if "hub" in which: case_hub_480x640()
