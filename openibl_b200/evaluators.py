"""Host-side mirror of ibl/evaluators.py (reference lines cited per function).

Same call signatures and return types, different data flow:
  * extract_features keeps descriptors on the GPU for the whole loop (one D2H at the end, not one
    per batch, evaluators.py:58) and gathers across ranks with one all_gather;
  * Evaluator.evaluate never builds the [m,n] distance matrix: every rank ranks all queries
    against its own contiguous database slice with the fused distance+top-k kernel, the per-shard
    top-k candidates are all-gathered (NCCL over NVLink) and merged;
  * pairwise_distance still returns the dense matrix for the training callers that need a full
    argsort (netvlad_img.py:78), computed on the GPU.
"""
from __future__ import annotations

import time
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

from .engine import Engine

__all__ = ["extract_cnn_feature", "extract_features", "pairwise_distance", "spatial_nms",
           "evaluate_all", "recalls_from_topk", "sharded_topk", "Evaluator"]


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _to_torch(x):
    if torch.is_tensor(x):
        return x
    if type(x).__module__ == "numpy":
        return torch.from_numpy(x)
    raise ValueError("Cannot convert {} to torch tensor".format(type(x)))


def extract_cnn_feature(model, inputs, vlad=True, gpu=None):
    """evaluators.py:22-34: forward, pick vlad/pooled output, L2 (idempotent for vlad)."""
    model.eval()
    inputs = _to_torch(inputs).cuda(gpu, non_blocking=True)
    with torch.no_grad():
        outputs = model(inputs)
    if isinstance(outputs, (list, tuple)):
        x_pool, x_vlad = outputs
        outputs = x_vlad if vlad else x_pool
    return Engine.get(outputs.device).l2_normalize_rows(outputs)


def extract_features(model, data_loader, dataset, print_freq=10, vlad=True, pca=None, gpu=None,
                     sync_gather=False):
    """evaluators.py:36-103 -> OrderedDict{fname: CPU FloatTensor[D]} in `dataset` order.
    The loader yields (imgs, fnames, pids, x, y); rank r holds the r-th contiguous padded slice
    (DistributedSliceSampler, sampler.py:194-223)."""
    model.eval()
    rank, world = _rank_world()
    if pca is not None:
        pca.load(gpu=gpu)
    feats = []
    end = time.time()
    bt_sum = 0.0
    with torch.no_grad():
        for i, (imgs, fnames, _, _, _) in enumerate(data_loader):
            out = extract_cnn_feature(model, imgs, vlad, gpu=gpu)
            if pca is not None:
                out = pca.infer(out)
            feats.append(out)            # stays on the GPU; no per-batch sync
            bt = time.time() - end
            bt_sum += bt
            end = time.time()
            if (i + 1) % print_freq == 0 and rank == 0:
                print("Extract Features: [{}/{}]\tTime {:.3f} ({:.3f})".format(
                    i + 1, len(data_loader), bt, bt_sum / (i + 1)))
    local = torch.cat(feats) if feats else torch.empty(0, 0, device="cuda")
    if world > 1:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        allf = torch.cat(parts)
    else:
        allf = local
    allf = allf[: len(dataset)].cpu()      # drop the sampler's wrap-around padding
    features = OrderedDict()
    for item, row in zip(dataset, allf):
        features[item[0]] = row
    return features


def pairwise_distance(features, query=None, gallery=None, metric=None):
    """evaluators.py:105-130 -> (dist_m [m,n] CPU tensor, x.numpy(), y.numpy()); the self-distance
    branch (:106-114) when both lists are None.  The GEMM runs on the GPU."""
    eng = Engine.get()
    dev = torch.device("cuda", eng.device)
    if query is None and gallery is None:
        x = torch.stack(list(features.values())).view(len(features), -1)
        if metric is not None:
            x = metric.transform(x)
        xg = x.to(dev)
        return eng.l2dist_self(xg).cpu(), None, None   # 2|x_i|^2 - 2 x_i.x_j, as the reference (:110-113)
    x = torch.stack([features[f] for f, _, _, _ in query]).view(len(query), -1)
    y = torch.stack([features[f] for f, _, _, _ in gallery]).view(len(gallery), -1)
    if metric is not None:
        x = metric.transform(x)
        y = metric.transform(y)
    d = eng.l2dist_dense(x.to(dev), y.to(dev))
    return d.cpu(), x.numpy(), y.numpy()


def spatial_nms(pred, db_ids, topN):
    """evaluators.py:132-140: first occurrence of each place id among the first topN."""
    kept, seen = [], set()
    for p in pred[:topN]:
        pid = db_ids[p]
        if pid not in seen:
            seen.add(pid)
            kept.append(p)
    return kept


def recalls_from_topk(topk_idx, gt, gallery, recall_topk=(1, 5, 10), nms=False):
    """The recall loop of evaluate_all (evaluators.py:151-162) on a [m,k] ranking."""
    topk_idx = np.asarray(topk_idx)
    db_ids = [db[1] for db in gallery]
    correct = np.zeros(len(recall_topk))
    for q, pred in enumerate(topk_idx):
        pred = pred[pred >= 0]
        if nms:
            seen, kept = set(), []
            for p in pred[: max(recall_topk) * 12]:
                pid = db_ids[p]
                if pid not in seen:
                    seen.add(pid)
                    kept.append(p)
            pred = np.asarray(kept, dtype=np.int64)
        for i, n in enumerate(recall_topk):
            if np.any(np.isin(pred[:n], gt[q])):
                correct[i:] += 1
                break
    return correct / len(gt)


def evaluate_all(distmat, gt, gallery, recall_topk=[1, 5, 10], nms=False):
    """evaluators.py:142-167 from a dense matrix; only the consumed ranks are selected (on the GPU)."""
    rank, _ = _rank_world()
    k = max(recall_topk) * (12 if nms else 1)
    d = _to_torch(distmat).float()
    eng = Engine.get()
    k = min(k, d.shape[1])
    if k > 1024:
        # the reference argsorts the whole row and so accepts any recall_topk; the selection kernel keeps
        # up to 1024 ranks per query (Tokyo nms needs 120) -- fail loudly rather than score an empty ranking
        raise NotImplementedError(
            f"evaluate_all needs the first {k} ranks per query; ibl_topk_rows supports at most 1024")
    d = d.to(torch.device("cuda", eng.device))
    _, order = eng.topk_rows(d, k)
    order = order.cpu().numpy()
    recalls = recalls_from_topk(order, gt, gallery, recall_topk, nms)
    if rank == 0:
        print("Recall Scores:")
        for i, kk in enumerate(recall_topk):
            print("  top-{:<4}{:12.1%}".format(kk, recalls[i]))
    return recalls


def sharded_topk(q: torch.Tensor, db_shard: torch.Tensor, k: int, idx_base: int, n_valid: int,
                 _rank_fn=None, _merge_fn=None):
    """Distance + top-k of all queries against this rank's database slice, ONE all-gather of the
    [m,k] candidates (dist and idx packed in one int64 tensor), k-way merge.  Returns
    (dist [m,k], idx [m,k]), identical on every rank.  `_rank_fn` / `_merge_fn` exist only so the
    gloo CPU test can drive the distributed plumbing without a GPU; the product path is the CUDA
    engine."""
    _, world = _rank_world()
    if _rank_fn is None:
        eng = Engine.get(q.device)
        _rank_fn = lambda qq, dd, kk, base, nv: eng.l2dist_topk(qq, dd, kk, idx_base=base, n_valid=nv)
        _merge_fn = eng.topk_merge
    cd, ci = _rank_fn(q, db_shard, k, idx_base, n_valid)
    if world == 1:
        return cd, ci
    packed = torch.stack([cd.contiguous().view(torch.int32).to(torch.int64), ci])   # [2,m,k] int64
    gathered = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(gathered, packed)
    gd = torch.stack([g[0].to(torch.int32).view(torch.float32) for g in gathered])
    gi = torch.stack([g[1] for g in gathered])
    return _merge_fn(gd, gi, k)


class Evaluator(object):
    """evaluators.py:170-201."""

    def __init__(self, model):
        self.model = model
        self.rank, _ = _rank_world()

    def evaluate(self, query_loader, dataset, query, gallery, ground_truth, gallery_loader=None,
                 vlad=True, pca=None, rerank=False, gpu=None, sync_gather=False, nms=False, rr_topk=25,
                 lambda_value=0):
        if gallery_loader is not None:
            features = extract_features(self.model, query_loader, query, vlad=vlad, pca=pca, gpu=gpu,
                                        sync_gather=sync_gather)
            features.update(extract_features(self.model, gallery_loader, gallery, vlad=vlad, pca=pca, gpu=gpu,
                                             sync_gather=sync_gather))
        else:
            features = extract_features(self.model, query_loader, dataset, vlad=vlad, pca=pca, gpu=gpu,
                                        sync_gather=sync_gather)
        rank, world = _rank_world()
        dev = torch.device("cuda", torch.cuda.current_device() if gpu is None else gpu)
        x = torch.stack([features[f] for f, _, _, _ in query]).to(dev)
        n = len(gallery)
        per = (n + world - 1) // world
        lo, hi = min(rank * per, n), min((rank + 1) * per, n)
        shard = torch.zeros(max(per, 1), x.shape[1], device=dev)
        if hi > lo:
            shard[: hi - lo] = torch.stack([features[f] for f, _, _, _ in gallery[lo:hi]]).to(dev)
        k = min(max(10 * (12 if nms else 1), 1), 128)
        _, idx = sharded_topk(x, shard, k, idx_base=lo, n_valid=hi - lo)
        recalls = recalls_from_topk(idx.cpu().numpy(), ground_truth, gallery, (1, 5, 10), nms)
        if self.rank == 0:
            print("Recall Scores:")
            for i, kk in enumerate((1, 5, 10)):
                print("  top-{:<4}{:12.1%}".format(kk, recalls[i]))
        if not rerank:
            return recalls
        # evaluators.py:194-201: k-reciprocal re-ranking needs the dense q-g, q-q and g-g matrices; they are
        # built on the GPU (tcgen05 dense distance tiles) and re-ranked there (utils/rerank.py), rank 0 only,
        # as in the reference (the other ranks score the original matrix).
        eng = Engine.get(dev)
        y = torch.stack([features[f] for f, _, _, _ in gallery]).to(dev)
        distmat = eng.l2dist_dense(x, y)
        if self.rank == 0:
            print("Applying re-ranking ...")
            from .utils.rerank import re_ranking
            distmat = re_ranking(distmat, eng.l2dist_dense(x, x), eng.l2dist_dense(y, y), k1=rr_topk, k2=1,
                                 lambda_value=lambda_value)
        return evaluate_all(distmat, ground_truth, gallery, nms=nms)
