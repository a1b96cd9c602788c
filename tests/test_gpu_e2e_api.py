"""End-to-end drive of the reference-shaped API (the flow of the reference's examples/test.py) under
torchrun, NCCL backend: init_dist -> DistributedSliceSampler loaders -> DDP(models.create(...)) ->
Evaluator.evaluate.  World size 1 always; world size 2 when the box has two GPUs.  The recalls must
equal the CPU oracle's on the same seeded images and weights."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

N_DB, N_Q, H, W = 40, 12, 64, 96


def _run(nproc, port, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "examples", "eval_synthetic.py"), "--launcher", "pytorch",
           "--n-db", str(N_DB), "--n-q", str(N_Q), "--height", str(H), "--width", str(W), "--test-batch-size", "5", *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    m = re.search(r"RECALLS ([0-9.,]+)", out.stdout)   # the line printed after the (optional) re-ranking
    assert m, out.stdout[-2000:]
    return np.array([float(v) for v in m.group(1).split(",")])


def _oracle_recalls(rerank=None):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from eval_synthetic import SeededImages
    from oracle import ibl_oracle as O
    from openibl_b200 import datasets, models
    ds = datasets.create("synthetic", None, n_db=N_DB, n_q=N_Q, seed=0)
    torch.manual_seed(0)
    base = models.create("vgg16", pretrained=False)
    pool = models.create("netvlad", dim=512)
    from openibl_b200 import synth
    p = synth.make_netvlad_params(seed=0, sharp=True)
    pool.centroids.data.copy_(p["centroids"])
    pool.conv.weight.data.copy_(p["conv_weight"])
    sd = models.create("embednet", base, pool).state_dict()

    def feats(items):
        data = SeededImages(items, H, W)
        x = torch.stack([data[i][0] for i in range(len(items))])
        with torch.no_grad():
            return O.extract_descriptor(x, sd, vlad=True)

    q, db = feats(ds.q_test), feats(ds.db_test)
    d = O.pairwise_distance(q, db).numpy()
    if rerank is not None:     # evaluators.py:194-199 with the host re-ranking (pinned by tests/golden/rerank.npz)
        from openibl_b200.utils.rerank import re_ranking
        d = re_ranking(d, O.pairwise_distance(q, q).numpy(), O.pairwise_distance(db, db).numpy(), k1=rerank[0], k2=1,
                       lambda_value=rerank[1])
    return O.evaluate_all(d, ds.test_pos, [p[1] for p in ds.db_test])


def test_evaluator_flow_matches_oracle_world1_and_world2():
    want = _oracle_recalls()
    got1 = _run(1, 29721)
    assert np.allclose(got1, want, atol=2e-6), (got1, want)     # the script prints 6 decimals
    if torch.cuda.device_count() >= 2:
        got2 = _run(2, 29722)
        assert np.allclose(got2, want, atol=2e-6), (got2, want)


def test_evaluator_rerank_flow_matches_oracle():
    """examples/test.py --rerank: Evaluator.evaluate(rerank=True, rr_topk, lambda_value) end to end."""
    want = _oracle_recalls(rerank=(10, 0.3))
    got = _run(1, 29723, extra=("--rerank", "--rr-topk", "10", "--lambda-value", "0.3"))
    assert np.allclose(got, want, atol=2e-6), (got, want)
