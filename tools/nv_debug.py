import os, sys
os.environ["IBL_NV_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openibl_b200 import synth
from openibl_b200.engine import Engine
eng = Engine.get(0)
p = synth.make_netvlad_params(seed=8, sharp=True)
feat = torch.randn(32, 30, 40, 512, device="cuda")
w, c = p["conv_weight"].cuda(), p["centroids"].cuda()
for i in range(3):
    eng.netvlad_forward(feat, w, c, nhwc=True, want_raw=False, want_norm=True)
torch.cuda.synchronize()
