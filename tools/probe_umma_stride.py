"""Probe (GPU): can a K-major SW128 UMMA operand start at a non-1024-aligned row and use a group stride that
is not a multiple of 1024 B?  Prints, per (s0, group_rows, base_mode), whether D matches the expected view."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openibl_b200.engine import Engine, _ptr, _stream
from openibl_b200._cabi import check

eng = Engine.get(0)
torch.manual_seed(0)
rows = 200
A = torch.randint(-8, 9, (rows, 64), device="cuda").to(torch.bfloat16)
B = torch.randint(-8, 9, (64, 64), device="cuda").to(torch.bfloat16)
D = torch.empty(128, 64, device="cuda")
m = torch.arange(128, device="cuda")
for group_rows in (8, 10, 18, 12):
    for s0 in (0, 1, 3, 8, 11, 21):
        if s0 + 15 * group_rows + 8 > rows:
            continue
        idx = s0 + (m // 8) * group_rows + (m % 8)
        want = A[idx].float() @ B.float().t()
        res = []
        for mode in (0, 1):
            D.zero_()
            check(eng.lib.ibl_debug_umma_strided(eng.h, _ptr(A), rows, _ptr(B), s0, group_rows, mode, _ptr(D), _stream(0)), "probe")
            torch.cuda.synchronize()
            res.append(bool(torch.equal(D, want)))
        print(f"group_rows {group_rows:2d}  s0 {s0:2d}   base_offset=0: {res[0]}   base_offset=phase: {res[1]}", flush=True)
