#!/usr/bin/env python
"""Golden vectors for one SFRS training step (BASELINE configs[4]): the UNMODIFIED reference `SFRSTrainer._forward`
(ibl/trainers.py:235-259) driving the UNMODIFIED reference `EmbedRegionNet` (ibl/models/netvlad.py:112-207) on CPU,
generation 0 (image-level SARE loss) and generation 1 (hard-region loss), loss = hard + 0.5 * soft, backward, with
the trunk frozen below conv5 as `train_layers='conv5'` does (vgg.py:50-53).  TEST INFRASTRUCTURE; build container
only (needs /root/reference).

The reference's `_forward_train` cannot view a tuple_size > 1 batch on torch 2.x (netvlad.py:192, .view of a
non-contiguous slice), so the B tuples are run one at a time with tuple_size = 1 and averaged -- which is exactly what
the batched formulas compute (mean over tuples of per-tuple losses, trainers.py:247-257).

    python oracle/gen_golden_sfrs.py     # writes tests/golden/sfrs_step.npz
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
from ibl.trainers import SFRSTrainer          # the reference's trainer
from openibl_b200 import synth

B, NEG, NDIFF, H, W = 2, 2, 2, 64, 96
TEMP = [0.07, 0.07]


def build(seed):
    sd = synth.make_state_dict(seed=seed, sharp=True, with_pca=False, bias_scale=0.02)
    m = ref_models.create("embedregionnet", ref_models.create("vgg16", pretrained=False),
                          ref_models.create("netvlad", dim=512), tuple_size=1)
    m.load_state_dict(sd)
    for layer in list(m.base_model.base.children())[:24]:      # what pretrained=True + train_layers='conv5' freezes
        for p in layer.parameters():
            p.requires_grad = False
    return m.train()


easy, diff = synth.make_sfrs_tuples(seed=31, tuples=B, neg_num=NEG, n_diff=NDIFF, height=H, width=W)
out = {}
for gen in (0, 1):
    model, cache = build(13), build(23)
    tr = SFRSTrainer(model, cache, margin=0.1, neg_num=NEG, gpu=None, temp=TEMP)
    hard, soft = 0.0, 0.0
    for t in range(B):
        lh, ls = tr._forward(easy[t:t + 1], diff[t:t + 1], "sare_ind", gen)
        ((lh + 0.5 * ls) / B).backward()
        hard += lh.item() / B
        soft += ls.item() / B
    out[f"g{gen}_loss_hard"], out[f"g{gen}_loss_soft"] = np.float64(hard), np.float64(soft)
    for slot in (24, 26, 28):
        conv = model.base_model.base[slot]
        out[f"g{gen}_w{slot}"] = conv.weight.grad.numpy()[::8, ::8].copy()
        out[f"g{gen}_b{slot}"] = conv.bias.grad.numpy().copy()
        out[f"g{gen}_w{slot}_norm"] = np.float64(conv.weight.grad.double().norm().item())
    out[f"g{gen}_centroids"] = model.net_vlad.centroids.grad.numpy()[:, ::4].copy()
    out[f"g{gen}_conv_w"] = model.net_vlad.conv.weight.grad.numpy()[:, ::4, 0, 0].copy()
    assert model.base_model.base[21].weight.grad is None
    print(f"gen {gen}: loss_hard {hard:.6f} loss_soft {soft:.6f}")
# loss algebra alone on random region descriptors (the step above always picks the global region as the hardest):
# reference _get_loss for the three loss types and _get_hard_loss with a non-trivial region choice
g = torch.Generator().manual_seed(5)
unit = lambda *shape: torch.nn.functional.normalize(torch.randn(*shape, generator=g), dim=-1)
tr = SFRSTrainer(None, None, margin=0.1, neg_num=3, gpu=None, temp=TEMP)
la, lp, ln = unit(4, 64), unit(4, 64), unit(4, 3, 64)
out["u_anchors"], out["u_positives"], out["u_negatives"] = la.numpy(), lp.numpy(), ln.numpy()
for lt in ("triplet", "sare_joint", "sare_ind"):
    out[f"u_loss_{lt}"] = np.float64(tr._get_loss(la, lp, ln, 4, lt).item())
hn, hs = unit(3, 9, 64), torch.randn(3, 9, generator=g)
out["u_hard_negatives"], out["u_hard_scores"] = hn.numpy(), hs.numpy()
out["u_hard_loss"] = np.float64(tr._get_hard_loss(la[0].contiguous(), lp[0].contiguous(), hn, hs, "sare_ind").item())
path = os.path.join(ROOT, "tests", "golden", "sfrs_step.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
