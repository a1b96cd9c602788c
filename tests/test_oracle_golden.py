"""Pins oracle/ibl_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/gen_golden.py in the build container)."""
import numpy as np
import torch

from conftest import load_golden, rel_l2
from oracle import ibl_oracle as O
from openibl_b200 import synth

TOL = 2e-6  # same torch CPU kernels, different op order in NetVLAD: fp32 rounding only


def test_small_all_stages():
    g = load_golden("small_96x128")
    sd = synth.make_state_dict(seed=5, with_pca=True, pca_dim=128, bias_scale=0.05)
    x = synth.make_images(seed=6, batch=2, height=96, width=128)
    feat = O.vgg16_trunk(x, sd)
    assert rel_l2(feat, g["feat"]) < TOL
    assert rel_l2(O.global_max_pool(feat), g["pool"]) < TOL
    raw = O.netvlad(feat, sd["net_vlad.conv.weight"], sd["net_vlad.centroids"])
    assert rel_l2(raw, g["raw_vlad"]) < 5e-6
    v = O.vlad_normalize(raw)
    assert rel_l2(v, g["vlad"]) < 5e-6
    d = O.pca_whiten(v, sd["pca_layer.weight"], sd["pca_layer.bias"])
    assert rel_l2(d, g["desc"]) < 5e-6
    assert rel_l2(O.embednetpca_forward(x, sd), g["desc"]) < 5e-6


def test_odd_sizes_floor_pooling():
    g = load_golden("odd_70x90")
    sd = synth.make_state_dict(seed=7, with_pca=False, bias_scale=0.05)
    x = synth.make_images(seed=8, batch=1, height=70, width=90)
    pool, v = O.embednet_forward(x, sd)
    assert tuple(g["feat"].shape) == (1, 512, 4, 5)
    assert rel_l2(pool, g["pool"]) < TOL
    assert rel_l2(v, g["vlad"]) < 5e-6


def test_netvlad_soft_and_sharp():
    g = load_golden("netvlad_unit")
    gen = torch.Generator().manual_seed(11)
    feat = torch.randn(2, 512, 30, 40, generator=gen) * 3.0 + 0.5
    for tag, sharp in (("soft", False), ("sharp", True)):
        p = synth.make_netvlad_params(seed=3, sharp=sharp)
        raw = O.netvlad(feat, p["conv_weight"], p["centroids"])
        assert rel_l2(raw, g[f"{tag}_raw"]) < 1e-5, tag
        assert rel_l2(O.vlad_normalize(raw), g[f"{tag}_vlad"]) < 1e-5, tag
    # the sharp case really is sharp (softmax needs max subtraction), the soft one is ~uniform
    assert g["sharp_maxprob"] > 0.5 and g["soft_maxprob"] < 0.05
    assert abs(synth.make_netvlad_params(seed=3, sharp=True)["alpha"] - g["alpha"]) < 1e-3 * g["alpha"]


def test_pca_unit():
    g = load_golden("pca_unit")
    p = synth.make_pca_params(seed=9, in_dim=32768, out_dim=64)
    gen = torch.Generator().manual_seed(12)
    v = torch.nn.functional.normalize(torch.randn(5, 32768, generator=gen), dim=1)
    assert rel_l2(O.pca_whiten(v, p["weight"], p["bias"]), g["out"]) < 5e-6


def test_retrieval_distance_ranking_recall():
    g = load_golden("retrieval")
    q, db, gt = synth.make_gallery(n_db=1500, n_q=300, dim=512, sigma=0.28)
    assert np.array_equal(gt.numpy(), g["gt"])
    d = O.pairwise_distance(q, db).numpy()
    assert np.abs(d[:32] - g["dist_sub"]).max() < 5e-6
    dist_k, idx_k = O.topk_from_distmat(d, 10)
    assert np.array_equal(idx_k, g["top10"])
    assert np.abs(dist_k - g["top10_dist"]).max() < 5e-6
    gt_list = [np.array([int(t)]) for t in gt]
    pids = [i // 3 for i in range(db.size(0))]
    assert np.array_equal(O.evaluate_all(d, gt_list, pids), g["recalls"])
    assert np.array_equal(O.evaluate_all(d, gt_list, pids, nms=True), g["recalls_nms"])
    feats = torch.cat([q, db])[:64]
    assert np.abs(O.self_distance(feats).numpy() - g["self_dist"]).max() < 5e-6


def test_hub_480x640_config1():
    """BASELINE config 1 (CPU plumbing): full-size image through the whole model."""
    g = load_golden("hub_480x640")
    sd = synth.make_state_dict(seed=0, with_pca=True)
    x = synth.make_images(seed=1, batch=1)
    with torch.no_grad():
        feat = O.vgg16_trunk(x, sd)
        assert rel_l2(feat[:, ::8, ::3, ::4], g["feat_sub"]) < TOL
        assert abs(feat.double().abs().sum().item() - g["feat_abs_sum"]) < 1e-6 * g["feat_abs_sum"]
        raw = O.netvlad(feat, sd["net_vlad.conv.weight"], sd["net_vlad.centroids"])
        assert rel_l2(raw[:, ::4, ::8], g["raw_vlad_sub"]) < 1e-5
        v = O.vlad_normalize(raw)
        assert rel_l2(v, g["vlad"]) < 1e-5
        assert rel_l2(O.global_max_pool(feat), g["pool"]) < TOL
        d = O.pca_whiten(v, sd["pca_layer.weight"], sd["pca_layer.bias"])
    assert d.shape == (1, 4096)
    assert abs(float(d.norm()) - 1.0) < 1e-5
    assert rel_l2(d, g["desc"]) < 1e-5
