#!/usr/bin/env python
"""Golden vectors for k-reciprocal re-ranking: the UNMODIFIED reference function ibl/utils/rerank.py:32-100
(`re_ranking`, as called by Evaluator.evaluate, evaluators.py:194-199) on seeded descriptor sets.
TEST INFRASTRUCTURE; build container only (needs /root/reference).

    python oracle/gen_golden_rerank.py     # writes tests/golden/rerank.npz
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("IBL_REFERENCE", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_rerank", os.path.join(REF, "ibl", "utils", "rerank.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def dist(x, y):   # the arithmetic of pairwise_distance (evaluators.py:127-129), fp32
    m, n = x.size(0), y.size(0)
    d = x.pow(2).sum(1, keepdim=True).expand(m, n) + y.pow(2).sum(1, keepdim=True).expand(n, m).t()
    return d.addmm(x, y.t(), beta=1, alpha=-2)


out = {}
cases = [("a", 40, 150, 24, 25, 1, 0.0), ("b", 30, 90, 16, 8, 1, 0.3), ("c", 25, 80, 16, 10, 3, 0.2)]
for name, m, n, d, k1, k2, lam in cases:
    g = torch.Generator().manual_seed(ord(name) + 5)
    centers = torch.randn(12, d, generator=g)
    db = torch.nn.functional.normalize(centers[torch.randint(0, 12, (n,), generator=g)] + 0.35 * torch.randn(n, d, generator=g), dim=1)
    q = torch.nn.functional.normalize(centers[torch.randint(0, 12, (m,), generator=g)] + 0.35 * torch.randn(m, d, generator=g), dim=1)
    qg, qq, gg = dist(q, db), dist(q, q), dist(db, db)
    final = ref.re_ranking(qg.numpy().copy(), qq.numpy().copy(), gg.numpy().copy(), k1=k1, k2=k2, lambda_value=lam)
    out.update({f"{name}_q": q.numpy(), f"{name}_db": db.numpy(), f"{name}_qg": qg.numpy(), f"{name}_qq": qq.numpy(),
                f"{name}_gg": gg.numpy(), f"{name}_final": final.astype(np.float32),
                f"{name}_params": np.array([k1, k2, lam], dtype=np.float64)})
path = os.path.join(ROOT, "tests", "golden", "rerank.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
