"""Checkpoint I/O mirrored from ibl/utils/serialization.py:14-81 (same file format and key rules)."""
import json
import os
import os.path as osp
import shutil

import torch
from torch.nn import Parameter


def mkdir_if_missing(d):
    os.makedirs(d, exist_ok=True)


def read_json(fpath):
    with open(fpath, "r") as f:
        return json.load(f)


def write_json(obj, fpath):
    mkdir_if_missing(osp.dirname(fpath))
    with open(fpath, "w") as f:
        json.dump(obj, f, indent=4, separators=(",", ": "))


def read_mat(path, key="dbStruct"):
    from scipy.io import loadmat
    return loadmat(path)[key].item()


def save_checkpoint(state, is_best, fpath="checkpoint.pth.tar"):
    mkdir_if_missing(osp.dirname(fpath))
    torch.save(state, fpath)
    if is_best:
        shutil.copy(fpath, osp.join(osp.dirname(fpath), "model_best.pth.tar"))


def load_checkpoint(fpath):
    if not osp.isfile(fpath):
        raise ValueError("=> No checkpoint found at '{}'".format(fpath))
    ckpt = torch.load(fpath, map_location=torch.device("cpu"), weights_only=False)
    print("=> Loaded checkpoint '{}'".format(fpath))
    return ckpt


def copy_state_dict(state_dict, model, strip=None, replace=None, add=None):
    """Copy by name and shape, skip mismatches, report what the model still misses
    (serialization.py:52-81)."""
    target = model.state_dict()
    copied = set()
    for name, param in state_dict.items():
        if strip is not None and replace is None and name.startswith(strip):
            name = name[len(strip):]
        if strip is not None and replace is not None:
            name = name.replace(strip, replace)
        if add is not None:
            name = add + name
        if name not in target:
            continue
        if isinstance(param, Parameter):
            param = param.data
        if param.size() != target[name].size():
            print("mismatch:", name, param.size(), target[name].size())
            continue
        target[name].copy_(param)
        copied.add(name)
    missing = set(target.keys()) - copied
    if missing:
        print("missing keys in state_dict:", missing)
    return model
