import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openibl_b200.engine import Engine, _ptr, _stream
from openibl_b200._cabi import check
eng = Engine.get(0)
torch.manual_seed(0)
A = torch.randn(128, 128, device="cuda"); B = torch.randn(128, 64, device="cuda")
C = torch.zeros(128, 64, device="cuda")
check(eng.lib.ibl_debug_gemm_tn(eng.h, _ptr(A), _ptr(B), _ptr(C), _stream(0)), "tn")
ref = A.double().t() @ B.double()
err = ((C.double() - ref).norm() / ref.norm()).item()
print("MN-major TN GEMM rel err", err, "PASS" if err < 1e-4 else "FAIL")
if err >= 1e-4:
    # diagnostics: which structure does the result resemble?
    for name, cand in (("A B (no transpose)", A.double() @ B.double()[:128] if False else None),):
        pass
    print(C[:4, :8]); print(ref[:4, :8].float())
