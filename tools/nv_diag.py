import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openibl_b200 import synth
from openibl_b200.engine import Engine
eng = Engine.get(0)
p = synth.make_netvlad_params(seed=8, sharp=True)
w, c = p["conv_weight"].cuda(), p["centroids"].cuda()
for B in (5, 32, 40):
    torch.manual_seed(B)
    feat = torch.randn(B, 30, 40, 512, device="cuda")
    raw2, _ = eng.netvlad_forward(feat.permute(0, 3, 1, 2).contiguous(), w, c, nhwc=False, want_raw=True, want_norm=True)
    bad = 0
    for rep in range(200):
        raw, _ = eng.netvlad_forward(feat, w, c, nhwc=True, want_raw=True, want_norm=True)
        if rep % 10 == 0:
            torch.cuda.synchronize()
        if float((raw - raw2).norm() / raw2.norm()) > 1e-4:
            bad += 1
    print(f"B={B}: {bad}/200 bad", flush=True)
