import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import numpy as np, torch
from eval_synthetic import SeededImages
from oracle import ibl_oracle as O
from openibl_b200 import datasets, models, synth
from openibl_b200.engine import Engine
N_DB, N_Q, H, W = 40, 12, 64, 96
ds = datasets.create("synthetic", None, n_db=N_DB, n_q=N_Q, seed=0)
torch.manual_seed(0)
base = models.create("vgg16", pretrained=False); pool = models.create("netvlad", dim=512)
p = synth.make_netvlad_params(seed=0, sharp=True)
pool.centroids.data.copy_(p["centroids"]); pool.conv.weight.data.copy_(p["conv_weight"])
model = models.create("embednet", base, pool)
sd = {k: v.clone() for k, v in model.state_dict().items()}
def imgs(items):
    data = SeededImages(items, H, W)
    return torch.stack([data[i][0] for i in range(len(items))])
xq, xdb = imgs(ds.q_test), imgs(ds.db_test)
with torch.no_grad():
    oq, odb = O.extract_descriptor(xq, sd), O.extract_descriptor(xdb, sd)
model = model.cuda().eval()
eng = Engine.get(0)
for bs in (5, 40):
    with torch.no_grad():
        gq = torch.cat([model(xq[i:i+bs].cuda())[1] for i in range(0, N_Q, bs)]).cpu()
        gdb = torch.cat([model(xdb[i:i+bs].cuda())[1] for i in range(0, N_DB, bs)]).cpu()
    rel = lambda a, b: ((a.double()-b.double()).norm(dim=1)/b.double().norm(dim=1)).max().item()
    print("batch", bs, "max rel err q", rel(gq, oq), "db", rel(gdb, odb))
d = O.pairwise_distance(oq, odb).numpy()
wd, wi = O.topk_from_distmat(d, 10)
dk, ik = eng.l2dist_topk(gq.cuda(), gdb.cuda(), 10)
print("top1 oracle", wi[:, 0], "\ntop1 gpu   ", ik[:, 0].cpu().numpy())
print("oracle recalls", O.evaluate_all(d, ds.test_pos, [p_[1] for p_ in ds.db_test]))
from openibl_b200.evaluators import recalls_from_topk
print("gpu recalls", recalls_from_topk(ik.cpu().numpy(), ds.test_pos, ds.db_test))
