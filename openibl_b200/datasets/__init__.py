"""Dataset registry (ibl/datasets/__init__.py:18-31).

Parsing the real Pittsburgh / Tokyo 24/7 .mat files is outside the accelerated path (SURVEY 2:
"OUT OF SCOPE -- real-dataset parsing"); the registry keeps the reference's names so that
examples/test.py's argparse (`choices=datasets.names()`) works, and ships the synthetic gallery
used by the parity tests and benchmarks."""
from .synthetic import SyntheticGallery

_factory = {"synthetic": SyntheticGallery}
_unported = ("pitts", "tokyo")


def names():
    return sorted(list(_factory.keys()) + list(_unported))


def create(name, root, *args, **kwargs):
    if name in _unported:
        raise NotImplementedError(
            f"dataset '{name}': .mat parsing of the real datasets is not part of the B200 hot path; "
            "use the reference's ibl.datasets for it, or 'synthetic'")
    if name not in _factory:
        raise KeyError("Unknown dataset:", name)
    return _factory[name](root, *args, **kwargs)
