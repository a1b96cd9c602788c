#!/usr/bin/env python
"""Golden vectors for the training tuple samplers: the UNMODIFIED reference DistributedRandomTupleSampler and
DistributedRandomDiffTupleSampler (ibl/utils/data/sampler.py:15-192) on a seeded synthetic distance matrix, two
"epochs" (the second uses the cached hard negatives), two ranks.  TEST INFRASTRUCTURE; build container only.

    python oracle/gen_golden_sampler.py     # writes tests/golden/sampler.npz
"""
import os, sys, types, random, warnings
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, os.environ.get("IBL_REFERENCE", "/root/reference"))
warnings.filterwarnings("ignore")
from ibl.utils.data.sampler import DistributedRandomTupleSampler, DistributedRandomDiffTupleSampler   # the reference's

NQ, NG, POS, NEGX = 37, 150, 6, 14
rng = np.random.RandomState(4)
dist = rng.rand(NQ, NG).astype(np.float32)
dist[:, 10] = dist[:, 11]                                   # exact ties: argsort must order them by index
jac = rng.rand(NQ, NG).astype(np.float32)
pos = np.stack([rng.choice(NG, POS, replace=False) for _ in range(NQ)])
neg = np.concatenate([pos, np.stack([rng.choice(NG, NEGX, replace=False) for _ in range(NQ)])], axis=1)
q = [("q%03d" % i, i, 0.0, 0.0) for i in range(NQ)]
g = [("g%03d" % i, 1000 + i, 0.0, 0.0) for i in range(NG)]
pos_l, neg_l = [p.tolist() for p in pos], [sorted(set(n.tolist())) for n in neg]
out = dict(dist=dist, jac=jac, pos=pos, neg=neg, sort_idx=torch.argsort(torch.from_numpy(dist), dim=1, stable=True).numpy())
sub = list(range(3, NQ, 2))
for name, cls, kw in (("tuple", DistributedRandomTupleSampler, dict(neg_num=4, neg_pool=30)),
                      ("diff", DistributedRandomDiffTupleSampler, dict(pos_num=3, pos_pool=5, neg_num=4, neg_pool=30))):
    for rank in (0, 1):
        s = cls(q, g, pos_l, neg_l, num_replicas=2, rank=rank, **kw)
        random.seed(11 + rank)
        if name == "tuple":
            s.sort_gallery(torch.from_numpy(dist), sub)
        else:
            s.sort_gallery(torch.from_numpy(dist), torch.from_numpy(jac), sub)
        s.sort_idx = torch.from_numpy(out["sort_idx"])        # make the tie order explicit (stable)
        for ep in (0, 1):
            rows = list(iter(s))                              # ragged for the diff sampler: pad with -1
            width = 2 + 4 + 3
            out[f"{name}_r{rank}_e{ep}"] = np.asarray([r + [-1] * (width - len(r)) for r in rows], dtype=np.int64)
        out[f"{name}_len"] = np.int64(len(s))
path = os.path.join(ROOT, "tests", "golden", "sampler.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB", out["tuple_r0_e0"].shape, out["diff_r1_e1"].shape)
