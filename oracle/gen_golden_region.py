#!/usr/bin/env python
"""Golden vectors for the SFRS region branch: the UNMODIFIED reference EmbedRegionNet in train mode
(ibl/models/netvlad.py:123-207) on CPU -- region similarity scores, region descriptors and the gradients of a
scalar loss w.r.t. the NetVLAD parameters.  TEST INFRASTRUCTURE; build container only (needs /root/reference).

    python oracle/gen_golden_region.py     # writes tests/golden/region_train.npz
"""
import os, sys, types, warnings
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, os.environ.get("IBL_REFERENCE", "/root/reference"))
warnings.filterwarnings("ignore")
from ibl import models as ref_models          # the reference's ibl
from openibl_b200 import synth

T, NIMG, H, W = 1, 5, 64, 96                  # 1 tuple x (1 anchor + 4 pairs); the reference's .view fails for T>1 on torch 2.x (netvlad.py:192)
sd = synth.make_state_dict(seed=13, sharp=True, with_pca=False, bias_scale=0.02)
base = ref_models.create("vgg16", pretrained=False)
pool = ref_models.create("netvlad", dim=512)
model = ref_models.create("embedregionnet", base, pool, tuple_size=T)
model.load_state_dict(sd)
model.train()
x = synth.make_images(seed=14, batch=T * NIMG, height=H, width=W)
score, va, vb = model(x)
g = torch.Generator().manual_seed(15)
wgt = torch.randn(score.shape, generator=g)
loss = (score * wgt).sum()
loss.backward()
out = dict(score=score.detach().numpy(), vlad_a=va.detach().numpy()[:, :, :, ::16], vlad_b=vb.detach().numpy()[:, :, :, ::16],
           loss=np.float64(loss.item()), grad_centroids=model.net_vlad.centroids.grad.numpy(),
           grad_conv_w=model.net_vlad.conv.weight.grad.numpy(), loss_weights=wgt.numpy())
path = os.path.join(ROOT, "tests", "golden", "region_train.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB", "score", tuple(score.shape), "va", tuple(va.shape), "vb", tuple(vb.shape))
