"""CPU oracle for the OpenIBL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file.  Nothing under openibl_b200/ or ibl/
imports it: the product path is the CUDA library and fails loudly without it.

What it is: a plain fp32 (optionally fp64) restatement, on the host, of the
reference's arithmetic for the path VGG16 trunk -> NetVLAD -> intra-norm/L2
-> PCA-whiten/L2 -> pairwise L2 distance -> ranking -> recall.  The reference
is itself a PyTorch program whose arithmetic lives in torch ops (SURVEY 8c:
no native code, no golden vectors of its own), so the restatement uses the
same stock torch CPU ops in functional form plus numpy for ranking.

Pinning: the reference itself is importable in the build container (with an
empty `h5py` stub).  oracle/gen_golden.py runs the *unmodified* reference
modules on seeded inputs and commits the outputs under tests/golden/;
tests/test_oracle_golden.py checks every function here against those vectors.
Parity is therefore pinned by reference outputs generated here (torch 2.11
CPU), not by vectors the reference ships (it ships none).

Every function cites the reference file:line it follows (paths relative to
the reference root).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from openibl_b200.synth import VGG16_PLAN


# --------------------------------------------------------------------------
# stage (i): VGG16 conv1_1..conv5_3
# --------------------------------------------------------------------------
def vgg16_trunk(x: torch.Tensor, sd: dict, prefix: str = "base_model.base.") -> torch.Tensor:
    """ibl/models/vgg.py:40-42,61-62: torchvision vgg16 `features[:-2]`:
    13x (conv3x3 s1 p1 + bias), ReLU after all but the last conv, MaxPool2x2
    after blocks 1-4.  x [B,3,H,W] -> [B,512,H/16,W/16]."""
    last_slot = [p for p in VGG16_PLAN if p != "P"][-1][0]
    for item in VGG16_PLAN:
        if item == "P":
            x = F.max_pool2d(x, kernel_size=2, stride=2)
            continue
        slot = item[0]
        w = sd[f"{prefix}{slot}.weight"].to(x.dtype)
        b = sd[f"{prefix}{slot}.bias"].to(x.dtype)
        x = F.conv2d(x, w, b, stride=1, padding=1)
        if slot != last_slot:
            x = F.relu(x)
    return x


def global_max_pool(feat: torch.Tensor) -> torch.Tensor:
    """ibl/models/vgg.py:67-68: AdaptiveMaxPool2d(1) + view -> [B,512]."""
    return feat.amax(dim=(2, 3))


# --------------------------------------------------------------------------
# stage (ii): NetVLAD + normalisations
# --------------------------------------------------------------------------
def netvlad(feat: torch.Tensor, conv_weight: torch.Tensor, centroids: torch.Tensor,
            normalize_input: bool = True) -> torch.Tensor:
    """ibl/models/netvlad.py:44-61.  feat [N,C,h,w] -> raw vlad [N,K,C].

    Written as the algebraically equal  sum_s a*x - cent*sum_s a  so that the
    [N,K,C,S] temporary of the reference (netvlad.py:56-59) is not built; the
    golden test pins this against the reference's literal formulation."""
    N, C = feat.shape[:2]
    K = centroids.shape[0]
    x = feat
    if normalize_input:
        x = F.normalize(x, p=2, dim=1)                       # netvlad.py:47
    xf = x.reshape(N, C, -1)                                 # [N,C,S]
    logits = torch.einsum("kc,ncs->nks", conv_weight.reshape(K, C).to(x.dtype), xf)  # :50
    a = torch.softmax(logits, dim=1)                         # :51
    vlad = torch.einsum("nks,ncs->nkc", a, xf)               # sum_s a * x
    vlad = vlad - centroids.to(x.dtype).unsqueeze(0) * a.sum(dim=2).unsqueeze(2)
    return vlad


def vlad_normalize(vlad: torch.Tensor) -> torch.Tensor:
    """ibl/models/netvlad.py:78-80 (== :100-102, :202-204): intra-normalise each
    cluster row, flatten k-major, global L2.  [N,K,C] -> [N,K*C]."""
    v = F.normalize(vlad, p=2, dim=2)
    v = v.reshape(v.shape[0], -1)
    return F.normalize(v, p=2, dim=1)


def pca_whiten(v: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """ibl/models/netvlad.py:105-108 == ibl/pca.py:117-121: 1x1 conv (a GEMM)
    + bias, then L2.  v [N,D], weight [P,D,1,1] -> [N,P]."""
    P = weight.shape[0]
    y = v @ weight.reshape(P, -1).to(v.dtype).t() + bias.to(v.dtype)
    return F.normalize(y, p=2, dim=-1)


def pca_load(U: np.ndarray, lams: np.ndarray, mu: np.ndarray, n_components: int = 4096,
             whitening: bool = True):
    """ibl/pca.py:96-106 without h5py: (U, lams, mu) -> conv weight [P,D,1,1], bias [P]."""
    U = U[:, :n_components]
    lams = lams[:n_components]
    if whitening:
        U = np.matmul(U, np.diag(1.0 / np.sqrt(lams)))
    Utmu = np.matmul(U.T, mu)
    weight = torch.from_numpy(np.ascontiguousarray(U.T)).float().reshape(n_components, -1, 1, 1)
    bias = torch.from_numpy(-Utmu).reshape(-1).float()
    return weight, bias


def pca_train(x: torch.Tensor, n_components: int = 4096):
    """ibl/pca.py:28-67 with torch.linalg.eigh in place of the removed torch.symeig
    (same ascending-eigenvalue contract).  x [N,dim] -> (U, lams, mu, Utmu) numpy."""
    x = x.t()
    n_pts, n_dims = x.size(1), x.size(0)
    mu = x.mean(1).unsqueeze(1)
    x = x - mu
    if n_dims <= n_pts:
        dual = False
        x2 = torch.matmul(x, x.t()) / (n_pts - 1)
    else:
        dual = True
        x2 = torch.matmul(x.t(), x) / (n_pts - 1)
    L, U = torch.linalg.eigh(x2)
    if n_components < x2.size(0):
        k_idx = torch.argsort(L, descending=True)[:n_components]
        L = torch.index_select(L, 0, k_idx)
        U = torch.index_select(U, 1, k_idx)
    lams = L.clone()
    lams[lams < 1e-9] = 1e-9
    if dual:
        U = torch.matmul(x, torch.matmul(U, torch.diag(1.0 / torch.sqrt(lams)) / np.sqrt(n_pts - 1)))
    Utmu = torch.matmul(U.t(), mu)
    return U.numpy(), lams.numpy(), mu.numpy(), Utmu.numpy()


# --------------------------------------------------------------------------
# model-level forwards
# --------------------------------------------------------------------------
def embednet_forward(x: torch.Tensor, sd: dict):
    """ibl/models/netvlad.py:73-82: -> (pool_x [B,512], vlad [B,32768])."""
    feat = vgg16_trunk(x, sd)
    vl = netvlad(feat, sd["net_vlad.conv.weight"], sd["net_vlad.centroids"])
    return global_max_pool(feat), vlad_normalize(vl)


def embednetpca_forward(x: torch.Tensor, sd: dict) -> torch.Tensor:
    """ibl/models/netvlad.py:95-110: -> [B,4096]."""
    _, v = embednet_forward(x, sd)
    return pca_whiten(v, sd["pca_layer.weight"], sd["pca_layer.bias"])


def extract_descriptor(x: torch.Tensor, sd: dict, vlad: bool = True, pca=None) -> torch.Tensor:
    """ibl/evaluators.py:22-34,56-57 minus .cuda(): model forward, pick vlad or
    pooled output, (idempotent) L2, optional PCA.infer."""
    if "pca_layer.weight" in sd and pca is None:
        out = embednetpca_forward(x, sd)
        return F.normalize(out, p=2, dim=-1)
    pool_x, v = embednet_forward(x, sd)
    out = F.normalize(v if vlad else pool_x, p=2, dim=-1)
    if pca is not None:
        out = pca_whiten(out, pca[0], pca[1])
    return out


# --------------------------------------------------------------------------
# stage (iii): distance, ranking, recall
# --------------------------------------------------------------------------
def pairwise_distance(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """ibl/evaluators.py:121-129: ||x||^2 + ||y||^2 - 2 x y^T, fp32, [m,n]."""
    m, n = x.size(0), y.size(0)
    d = torch.pow(x, 2).sum(dim=1, keepdim=True).expand(m, n) + \
        torch.pow(y, 2).sum(dim=1, keepdim=True).expand(n, m).t()
    return torch.addmm(d, x, y.t(), beta=1, alpha=-2)


def self_distance(x: torch.Tensor) -> torch.Tensor:
    """ibl/evaluators.py:106-114 (query is None and gallery is None): 2||x_i||^2 - 2 x x^T."""
    n = x.size(0)
    d = torch.pow(x, 2).sum(dim=1, keepdim=True) * 2
    return d.expand(n, n) - 2 * torch.mm(x, x.t())


def topk_from_distmat(dist: np.ndarray, k: int):
    """Ranking consumed by ibl/evaluators.py:143 (np.argsort, full sort); only the
    first max(recall_topk) (x12 with nms) columns are ever read (:151-159).
    Tie rule made explicit: lowest index first (stable sort)."""
    idx = np.argsort(dist, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(dist, idx, axis=1), idx


def spatial_nms(pred, db_ids, top_n):
    """ibl/evaluators.py:132-140: keep the first occurrence of every place id
    among the first top_n predictions."""
    keep, seen = [], set()
    for p in pred[:top_n]:
        pid = db_ids[p]
        if pid not in seen:
            seen.add(pid)
            keep.append(p)
    return keep


def recalls_from_ranking(sort_idx, gt, gallery_pids=None, recall_topk=(1, 5, 10), nms=False):
    """ibl/evaluators.py:142-167: a query counts for every N >= the first N at which
    any of pred[:N] is a ground-truth positive."""
    correct = np.zeros(len(recall_topk))
    for q, pred in enumerate(sort_idx):
        pred = list(pred)
        if nms:
            pred = spatial_nms(pred, gallery_pids, max(recall_topk) * 12)
        g = set(np.asarray(gt[q]).reshape(-1).tolist())
        for i, n in enumerate(recall_topk):
            if any(p in g for p in pred[:n]):
                correct[i:] += 1
                break
    return correct / len(gt)


def evaluate_all(dist: np.ndarray, gt, gallery_pids=None, recall_topk=(1, 5, 10), nms=False):
    """ibl/evaluators.py:142-167 end to end from a dense distance matrix."""
    k = max(recall_topk) * (12 if nms else 1)
    _, idx = topk_from_distmat(np.asarray(dist), min(k, dist.shape[1]))
    return recalls_from_ranking(idx, gt, gallery_pids, recall_topk, nms)
