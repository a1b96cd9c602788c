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


def _gallery(extra, nproc=1, port=29731):
    import json
    cmd = [sys.executable]
    if nproc > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                "--master-port", str(port)]
    cmd += [os.path.join(ROOT, "tools", "bench_gallery.py"), "--n-db", "301", "--n-q", "45", "--height", "64",
            "--width", "96", "--batch", "8", *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_config4_gallery_ranking_is_world_size_independent():
    """BASELINE configs[3] correctness: the sharded extract + rank + candidate all-gather gives the SAME top-10
    indices and Recall@1/5/10 at world size 8 (all ranks played on this GPU, same slicing and batch tails as the
    real run), at world size 3 (ragged slices) and at world size 1, and agrees with an fp64 ranking of a query
    subset.  With two GPUs the real NCCL path is compared as well (evaluators.py:76-101,142-167 replaced)."""
    one = _gallery(["--emulate-world", "1"])
    assert one["topk_sane"] and one["exact_fp64_subset_agreement"] > 0.99, one
    assert 0.3 < one["recalls"][0] <= 1.0
    for w in (8, 3):
        r = _gallery(["--emulate-world", str(w)])
        assert r["topk_index_hash"] == one["topk_index_hash"], (w, r, one)
        assert r["recalls"] == one["recalls"] and r["exact_fp64_subset_agreement"] > 0.99
    plain = _gallery([])
    assert plain["topk_index_hash"] == one["topk_index_hash"] and plain["recalls"] == one["recalls"]
    if torch.cuda.device_count() >= 2:
        two = _gallery([], nproc=2)
        assert two["topk_index_hash"] == one["topk_index_hash"] and two["recalls"] == one["recalls"], (two, one)


def test_sfrs_step_under_ddp_matches_reference_loss():
    """BASELINE configs[4]: one SFRS step (trainers.py:235-259) under torch.distributed.run + DistributedDataParallel
    over NCCL, on every GPU of the box (8 on the scaling box, 1 here if there is one): rank 0's losses equal the
    unmodified reference's CPU losses on the same tuples (tests/golden/sfrs_step.npz), gradients are finite, and all
    ranks hold identical parameters after the optimizer step."""
    import json
    from conftest import load_golden
    g = load_golden("sfrs_step")
    nproc = max(1, min(8, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", "29751", os.path.join(ROOT, "examples", "sfrs_step_synthetic.py"),
           "--launcher", "pytorch", "--tuple-size", "2", "--neg-num", "2", "--diff-num", "2", "--height", "64", "--width", "96",
           "--generation", "0", "--seed", "31"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("SFRS_STEP ")][-1][len("SFRS_STEP "):])
    assert r["world"] == nproc and r["finite"] and r["params_identical_across_ranks"] and r["grad_norm_rank0"] > 0
    assert abs(r["loss_hard"] - float(g["g0_loss_hard"])) < 2e-4 and abs(r["loss_soft"] - float(g["g0_loss_soft"])) < 6e-4
    assert r["trainable_params"] == 3 * (512 * 512 * 9 + 512) + 2 * 64 * 512
