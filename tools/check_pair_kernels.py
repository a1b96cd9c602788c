#!/usr/bin/env python
"""Quick GPU sanity check of the SM-pair kernels (conv 256-wide pair tiles, single-pass distance, round-1 bf16x3
distance) against fp64 / exact references; exits non-zero on mismatch.  Run under a short `timeout` first in a GPU
session so that a barrier-protocol mistake costs a minute, not the session."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from openibl_b200.engine import Engine

eng = Engine.get(0)
g = torch.Generator().manual_seed(0)
ok = True
for (N, H, W, cin, cout) in ((3, 30, 40, 512, 512), (2, 17, 23, 256, 512)):
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    y = eng.debug_conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), b.cuda(), relu=False, pool=False, mode=1, bn=256).cpu()
    err = float((y.double() - ref).norm() / ref.norm())
    print("conv pair", (N, H, W, cin, cout), "rel-L2", err, flush=True)
    ok &= err < 2e-5
from openibl_b200 import synth
q, db, gt = synth.make_gallery(n_db=5000, n_q=600, dim=512, sigma=0.28)
exact = (q.double().unsqueeze(1) - db.double().unsqueeze(0)).pow(2).sum(-1) if False else None
d = (q.double() @ db.double().t()) * -2 + (q.double() ** 2).sum(1, keepdim=True) + (db.double() ** 2).sum(1)
wi = d.topk(10, largest=False).indices
dk, ik = eng.l2dist_topk(q.cuda(), db.cuda(), 10)
torch.cuda.synchronize()
agree = float((ik.cpu() == wi).float().mean())
print("dist single-pass agree", agree, "flagged", eng.dist_flagged(), flush=True)
ok &= agree > 0.999
print("PAIR_OK" if ok else "PAIR_FAIL", flush=True)
sys.exit(0 if ok else 1)
