"""Pillow-exact bilinear resize of uint8 image batches on the GPU (the first stage of the reference's
test transform, `T.Resize((height, width))` on a PIL image, ibl/utils/data/__init__.py:37-42).

`pil_bilinear_coeffs` rebuilds Pillow's coefficient tables (src/libImaging/Resample.c: precompute_coeffs +
normalize_coeffs_8bpc) on the host in float64, operation for operation; the two integer passes run in libiblb200
(csrc/resize.cu, `ibl_resize_bilinear_u8`).  The result equals `PIL.Image.resize(size, Image.BILINEAR)` bit for bit
(tests/test_gpu_parity.py::test_gpu_resize_matches_pillow_bit_exact; the table builder is pinned on the CPU against
Pillow through a numpy emulation of the passes, tests/test_host_cpu.py)."""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x: float) -> float:
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


@lru_cache(maxsize=64)
def pil_bilinear_coeffs(in_size: int, out_size: int):
    """-> (bounds int32 [out,2] = (first input sample, count), kk int32 [out,ksize] fixed-point coefficients, ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale                       # bilinear filter support = 1
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def resize_u8_reference(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """numpy emulation of the two integer passes (HWC uint8) -- what the CUDA kernels compute; used by the CPU test
    that pins `pil_bilinear_coeffs` against Pillow itself."""
    def one_pass(a, out_size):                    # along axis 0
        n = a.shape[0]
        if out_size == n:
            return a
        bounds, kk, _ = pil_bilinear_coeffs(n, out_size)
        out = np.empty((out_size,) + a.shape[1:], dtype=np.uint8)
        for o in range(out_size):
            lo, cnt = bounds[o]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :cnt].astype(np.int64), a[lo:lo + cnt].astype(np.int64), axes=(0, 0))
            out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        return out
    tmp = one_pass(np.ascontiguousarray(img.transpose(1, 0, 2)), out_w).transpose(1, 0, 2)   # horizontal first
    return one_pass(np.ascontiguousarray(tmp), out_h)


def resize_u8(x_u8_nhwc, out_h: int, out_w: int):
    """GPU: uint8 [N,H,W,3] CUDA tensor -> uint8 [N,out_h,out_w,3], bit-identical to Pillow's bilinear resize."""
    import torch
    from ...engine import Engine
    eng = Engine.get(x_u8_nhwc.device)
    return eng.resize_u8(x_u8_nhwc, out_h, out_w)
