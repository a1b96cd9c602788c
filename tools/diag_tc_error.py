"""Where does the tcgen05 bf16x3 conv error come from?  Single layer vs fp64, error and signed bias vs K."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openibl_b200.engine import Engine
eng = Engine.get(0)
torch.manual_seed(0)
for cin in (64, 128, 256, 512):
    for relu_in in (False, True):
        N, H, W, cout = 1, 32, 32, 64
        x = torch.randn(N, cin, H, W)
        if relu_in:
            x = x.relu()
        w = torch.randn(cout, cin, 3, 3) * (2.0 / (cin * 9)) ** 0.5
        b = torch.zeros(cout)
        ref = torch.nn.functional.conv2d(x.double(), w.double(), None, padding=1).permute(0, 2, 3, 1)
        # what an exact evaluation of the 3 split products would give
        def split(t):
            hi = t.bfloat16().float(); lo = (t - hi).bfloat16().float(); return hi.double(), lo.double()
        xh, xl = split(x); wh, wl = split(w)
        c = lambda a, bb: torch.nn.functional.conv2d(a, bb, None, padding=1)
        ref3 = (c(xh, wh) + c(xh, wl) + c(xl, wh)).permute(0, 2, 3, 1)
        xd = x.permute(0, 2, 3, 1).contiguous().cuda()
        out = {}
        for name, mode in (("simt", 0), ("tc", 1)):
            y = eng.debug_conv3x3(xd, w.cuda(), b.cuda(), relu=False, pool=False, mode=mode).cpu().double()
            err = (y - ref).norm() / ref.norm()
            err3 = (y - ref3).norm() / ref3.norm()
            bias = ((y - ref) * ref.sign()).mean() / ref.abs().mean()
            out[name] = (err.item(), err3.item(), bias.item())
        print(f"cin={cin:4d} K={cin*9:5d} relu_in={relu_in}: simt err={out['simt'][0]:.2e} bias={out['simt'][2]:+.2e} | "
              f"tc err={out['tc'][0]:.2e} (vs exact 3-term {out['tc'][1]:.2e}) bias={out['tc'][2]:+.2e} | "
              f"3-term model err={(ref3-ref).norm()/ref.norm():.2e}")
