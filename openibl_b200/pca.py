"""Host-side mirror of ibl/pca.py: PCA(pca_n_components, pca_whitening, pca_parameters_path) with
.train(x) / .load(gpu) / .infer(data) (reference pca.py:21-123).

infer() is stage (iii-a) of the hot path and runs in libiblb200 (ibl_pca_l2).  Parameters are
stored as {U, lams, mu, Utmu} like the reference's h5 file (pca.py:79-84); h5py is optional --
without it the same arrays go to '<path>.npz'."""
from __future__ import annotations

import os

import numpy as np
import torch

from .engine import Engine

try:  # the reference hard-depends on h5py (pca.py:10); it is absent from this image
    import h5py  # type: ignore
    _HAVE_H5 = hasattr(h5py, "File")
except Exception:  # pragma: no cover
    h5py = None
    _HAVE_H5 = False


class PCA:
    def __init__(self, pca_n_components=4096, pca_whitening=True, pca_parameters_path="./logs/pca_params.h5"):
        self.pca_n_components = pca_n_components
        self.pca_whitening = pca_whitening
        self.pca_parameters_path = pca_parameters_path
        self.weight = None
        self.bias = None

    # ---- storage -------------------------------------------------------------------------
    def _npz_path(self):
        return self.pca_parameters_path + ".npz"

    def _save(self, **arrs):
        if _HAVE_H5:
            with h5py.File(self.pca_parameters_path, "w") as f:
                for k, v in arrs.items():
                    f.create_dataset(k, data=v)
        else:
            np.savez(self._npz_path(), **arrs)
            open(self.pca_parameters_path, "ab").close()   # marker so `osp.isfile(path)` holds (test.py:111)

    def _read(self):
        if _HAVE_H5 and os.path.getsize(self.pca_parameters_path) > 0:
            with h5py.File(self.pca_parameters_path, "r") as f:
                return {k: f[k][...] for k in ("U", "lams", "mu", "Utmu")}
        return dict(np.load(self._npz_path()))

    # ---- reference API -------------------------------------------------------------------
    def train(self, x):
        """pca.py:28-84 (relja_PCA): covariance or dual eigen-decomposition of the centred descriptors.

        The matrix products -- the [dims,dims] covariance or the [pts,pts] dual Gram matrix (10k x 10k x 32768 for
        examples/test.py:108-121), the back-projection U = X V and U^T mu -- run on the engine's own fp32 GEMM kernel
        (ibl_gemm_nt); the symmetric eigen-decomposition is torch.linalg.eigh (cuSOLVER; torch.symeig, which the
        reference calls, no longer exists -- same ascending-eigenvalue contract).  GPU only, like the rest of the
        engine."""
        print("calculating PCA parameters...")
        if not torch.cuda.is_available():
            raise RuntimeError("PCA.train runs on the GPU engine (there is no CPU fallback)")
        from ._cabi import CONV_SIMT_FP32
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        eng = Engine.get(dev)
        rows = x.to(dev).float().contiguous()                 # [pts, dims]: one descriptor per row
        n_pts, n_dims = rows.shape
        mu_row = rows.mean(0, keepdim=True)
        rows = rows - mu_row                                  # centred, still [pts, dims]
        dual = n_dims > n_pts
        if dual:
            x2 = eng.gemm_nt(rows, rows, alpha=1.0 / (n_pts - 1), mode=CONV_SIMT_FP32)            # X^T X, [pts,pts]
        else:
            cols = rows.t().contiguous()                                                          # X, [dims,pts]
            x2 = eng.gemm_nt(cols, cols, alpha=1.0 / (n_pts - 1), mode=CONV_SIMT_FP32)            # X X^T, [dims,dims]
        L, U = torch.linalg.eigh(x2)
        if self.pca_n_components < x2.size(0):
            keep = torch.argsort(L, descending=True)[: self.pca_n_components]
            L, U = L[keep], U[:, keep]
        lams = L.clamp_min(1e-9)
        if dual:
            coef = (U / torch.sqrt(lams).unsqueeze(0) / np.sqrt(n_pts - 1)).t().contiguous()     # [P, pts]
            U = eng.gemm_nt(rows.t().contiguous(), coef, mode=CONV_SIMT_FP32)                     # X V, [dims, P]
        Utmu = eng.gemm_nt(U.t().contiguous(), mu_row.contiguous(), mode=CONV_SIMT_FP32)          # [P, 1]
        self._save(U=U.cpu().numpy(), lams=lams.cpu().numpy(), mu=mu_row.t().cpu().numpy(), Utmu=Utmu.cpu().numpy())

    def load(self, gpu=None):
        """pca.py:86-106: W = (U diag(lams^-1/2))^T as [P, D, 1, 1], b = -W mu, on the GPU."""
        p = self._read()
        U = p["U"][:, : self.pca_n_components]
        lams = p["lams"][: self.pca_n_components]
        if self.pca_whitening:
            U = U @ np.diag(1.0 / np.sqrt(lams))
        Utmu = U.T @ p["mu"]
        dev = torch.device("cuda", torch.cuda.current_device() if gpu is None else gpu)
        self.weight = torch.from_numpy(np.ascontiguousarray(U.T)).float().view(self.pca_n_components, -1, 1, 1).to(dev)
        self.bias = torch.from_numpy(-Utmu).view(-1).float().to(dev)

    def infer(self, data):
        """pca.py:108-123: 1x1 conv (GEMM) + bias + L2 -> [N, pca_n_components]."""
        out = Engine.get(data.device).pca_l2(data, self.weight, self.bias)
        assert out.size(1) == self.pca_n_components
        return out
