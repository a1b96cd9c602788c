#!/usr/bin/env python
"""Is the synthetic gallery discriminative at scale?  Recall, distance scale, guard flags and ranking time with and
without the centred PCA bias (openibl_b200.gallery.center_pca).  usage: tools/diag_gallery.py [n_db] [n_q] [centre 0|1]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openibl_b200 import gallery, synth
from openibl_b200.engine import Engine

n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
centre = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H, W, B = 480, 640, 32
eng = Engine.get(0)
dev = torch.device("cuda", 0)
sd = {k: v.to(dev) for k, v in synth.make_state_dict(seed=0, with_pca=True).items()}
slots = synth.VGG16_CONV_SLOTS
eng.set_vgg16([sd[f"base_model.base.{s}.weight"] for s in slots], [sd[f"base_model.base.{s}.bias"] for s in slots])
eng.set_netvlad(sd["net_vlad.conv.weight"], sd["net_vlad.centroids"])
eng.set_pca(sd["pca_layer.weight"], sd["pca_layer.bias"])
if centre:
    gallery.center_pca(eng, sd["pca_layer.weight"], H, W, B)
gallery.run(eng, 64, 16, H, W, B, check_exact=False)
r = gallery.run(eng, n_db, n_q, H, W, B, check_exact=True)
r["flagged_last_call"] = eng.dist_flagged()
r["centre"] = centre
r["noise"] = gallery.NOISE
r["mean_top10_dist"] = r["topk_dist_checksum"] / (n_q * 10)
print(json.dumps(r))
