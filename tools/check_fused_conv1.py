#!/usr/bin/env python
"""Quick GPU check of the fused conv1_1+conv1_2 kernel against the two separate kernels (run in two processes by
tools/sessions/gpu_session3b.sh: the env switch is read once per process).  Prints a checksum line; `--time` also times it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from openibl_b200 import synth
from openibl_b200.engine import Engine, _ptr, check

eng = Engine.get(0)
sd = synth.make_vgg_weights(3, 0.05)
slots = synth.VGG16_CONV_SLOTS
eng.set_vgg16([sd[f"base.{s}.weight"].cuda() for s in slots], [sd[f"base.{s}.bias"].cuda() for s in slots])
out = []
for (n, h, w) in ((2, 96, 128), (1, 70, 90), (3, 32, 48), (1, 480, 640)):
    x = synth.make_images(seed=h, batch=n, height=h, width=w).cuda()
    a = eng.vgg16_prefix_forward(x, 3)          # conv1_1, conv1_2 (+pool), conv2_1  -> NHWC fp32
    torch.cuda.synchronize()
    out.append(a.double().cpu())
    print("shape", tuple(a.shape), "sum", float(a.double().sum()), "absmax", float(a.abs().max()), flush=True)
torch.save(out, sys.argv[1])
if "--time" in sys.argv:
    x = synth.make_images(seed=1, batch=32).cuda()
    for layer in ((13,) if os.environ.get("IBL_CONV1_FUSED", "1") != "0" else (0,)):
        ms = ctypes.c_float()
        check(eng.lib.ibl_debug_time_layer(eng.h, layer, _ptr(x), 32, 480, 640, 0, 5, ctypes.byref(ms)), "time")
        print("layer", layer, "ms", ms.value, flush=True)
