"""Loader helpers mirrored from ibl/utils/data/__init__.py:8-42."""
from .preprocessor import Preprocessor  # noqa: F401
from . import sampler  # noqa: F401

_MEAN = [0.48501960784313836, 0.4579568627450961, 0.4076039215686255]
_STD = [0.00392156862745098] * 3


class IterLoader:
    def __init__(self, loader, length=None):
        self.loader, self.length, self.iter = loader, length, None

    def __len__(self):
        return self.length if self.length is not None else len(self.loader)

    def new_epoch(self):
        self.iter = iter(self.loader)

    def next(self):
        try:
            return next(self.iter)
        except Exception:
            self.iter = iter(self.loader)
            return next(self.iter)


def get_transformer_train(height, width):
    import torchvision.transforms as T
    return T.Compose([T.ColorJitter(0.7, 0.7, 0.7, 0.5), T.Resize((height, width)), T.ToTensor(),
                      T.Normalize(mean=_MEAN, std=_STD)])


def get_transformer_test(height, width, tokyo=False):
    import torchvision.transforms as T
    return T.Compose([T.Resize(max(height, width) if tokyo else (height, width)), T.ToTensor(),
                      T.Normalize(mean=_MEAN, std=_STD)])
