"""DistributedSliceSampler (ibl/utils/data/sampler.py:194-223): rank r reads the r-th contiguous
slice of ceil(N/world) items, the tail wrapping to the head; extract_features un-pads by
truncation.  Part of the hot-path contract."""
import math

import torch.distributed as dist
from torch.utils.data.sampler import Sampler


def slice_bounds(n, world, rank):
    """(first, count, padded_count) of rank's slice."""
    per = int(math.ceil(n * 1.0 / world))
    lo = min(rank * per, n)
    return lo, max(0, min(n, lo + per) - lo), per


class DistributedSliceSampler(Sampler):
    def __init__(self, dataset, num_replicas=None, rank=None):
        if num_replicas is None:
            if not dist.is_available():
                raise RuntimeError("Requires distributed package to be available")
            num_replicas = dist.get_world_size()
        if rank is None:
            if not dist.is_available():
                raise RuntimeError("Requires distributed package to be available")
            rank = dist.get_rank()
        self.dataset, self.num_replicas, self.rank = dataset, num_replicas, rank
        n = len(dataset)
        self.num_samples = int(math.ceil(n * 1.0 / num_replicas))
        self.total_size = self.num_samples * num_replicas

    def __iter__(self):
        n = len(self.dataset)
        start = self.rank * self.num_samples
        return iter([(start + i) % n if start + i >= n else start + i for i in range(self.num_samples)])

    def __len__(self):
        return self.num_samples
