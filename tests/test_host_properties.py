"""Property tests of the host-side logic (CPU, no GPU): randomised inputs, checked against the oracle or against
invariants of the reference semantics."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import ibl_oracle as O


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 400), world=st.integers(1, 9))
def test_slice_bounds_cover_the_list_like_the_reference_sampler(n, world):
    """sampler.py:208-219: ceil(n/W) per rank, contiguous, the tail wraps to the head; un-padding keeps [:n]."""
    from openibl_b200.utils.data.sampler import DistributedSliceSampler, slice_bounds
    per = -(-n // world)
    seen = []
    for r in range(world):
        lo, cnt, p = slice_bounds(n, world, r)
        assert p == per and 0 <= cnt <= per and (cnt == 0 or lo + cnt <= n)
        idx = list(DistributedSliceSampler(list(range(n)), num_replicas=world, rank=r))
        assert len(idx) == per
        assert idx[:cnt] == list(range(lo, lo + cnt))          # the un-padded part is the contiguous slice
        ref = (list(range(n)) + list(range(n)) * world)[r * per:(r + 1) * per]   # padded list, reference order
        assert idx == ref
        seen += idx[:cnt]
    assert sorted(seen) == list(range(n))


@settings(max_examples=25, deadline=None)
@given(m=st.integers(1, 12), n=st.integers(12, 60), seed=st.integers(0, 10**6), nms=st.booleans())
def test_recalls_from_topk_equals_oracle_full_sort(m, n, seed, nms):
    """evaluate_all (evaluators.py:142-167) only consumes the first 10 (120 with nms) ranks: recalls from the
    truncated ranking equal recalls from the full argsort."""
    from openibl_b200.evaluators import recalls_from_topk
    rng = np.random.RandomState(seed)
    d = rng.permutation(m * n).reshape(m, n).astype(np.float32)        # distinct distances: no tie ambiguity
    gt = [rng.choice(n, size=rng.randint(1, 4), replace=False) for _ in range(m)]
    gallery = [("d%03d" % i, int(rng.randint(0, max(2, n // 3))), 0.0, 0.0) for i in range(n)]
    full = np.argsort(d, axis=1)
    want = O.recalls_from_ranking(full, gt, [g[1] for g in gallery], nms=nms)
    k = min(n, 120 if nms else 10)
    got = recalls_from_topk(full[:, :k], gt, gallery, nms=nms)
    assert np.array_equal(got, want)


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 10**6), k1=st.integers(3, 9), lam=st.sampled_from([0.0, 0.3, 1.0]))
def test_rerank_is_equivariant_to_gallery_order_and_bounded(seed, k1, lam):
    """k-reciprocal re-ranking depends on the gallery only through distances: permuting the gallery permutes
    the output columns; Jaccard distances lie in [0, 1]; lambda = 1 returns the normalised original distances."""
    from openibl_b200.utils.rerank import re_ranking
    g = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(9, 16, generator=g), dim=1)
    db = torch.nn.functional.normalize(torch.randn(31, 16, generator=g), dim=1)
    perm = torch.randperm(31, generator=g)

    def run(dbx):
        return re_ranking(O.pairwise_distance(q, dbx), O.pairwise_distance(q, q), O.pairwise_distance(dbx, dbx), k1=k1, k2=1,
                          lambda_value=lam)

    a, b = run(db), run(db[perm])
    assert torch.allclose(a[:, perm], b, atol=1e-6)
    assert float(a.min()) >= -1e-6 and float(a.max()) <= 1 + 1e-6
    if lam == 1.0:
        full = torch.cat([torch.cat([O.pairwise_distance(q, q), O.pairwise_distance(q, db)], 1),
                          torch.cat([O.pairwise_distance(q, db).t(), O.pairwise_distance(db, db)], 1)], 0).pow(2)
        norm = (full / full.max(dim=0, keepdim=True).values).t()
        assert torch.allclose(a, norm[:9, 9:], atol=1e-6)
