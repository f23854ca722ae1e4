"""GMM-based voice conversion -- drop-in for ``nnmnkwii.baseline.gmm`` (baseline/gmm.py:47-247).

SURVEY.md section 8f row 1: the in-repo caller that turns a joint source/target GMM and a source
utterance into the per-frame means ``E`` (Eq. 22) and diagonal variances ``D`` (Eq. 23) that
``paramgen.mlpg`` consumes.  The reference walks the frames in Python (one ``np.linalg.solve`` per
frame); here the mixture posteriors, the per-frame affine maps and the variances are evaluated for a
whole utterance -- or a whole batch of utterances -- on the GPU in float64 and handed to the MLPG
kernels without leaving the device.

The per-mixture matrices ``A[m] = covarYX[m] covarXX[m]^-1`` are formed once (the reference re-solves
per frame, gmm.py:113-115, 231-233).  Posteriors, arg-max mixture, affine map and variances run in the
float64 kernels of csrc/nnk_gmm.cu (C ABI ``nnk_gmm_logprob`` / ``nnk_gmm_map``): no (frames, mixtures,
dim) temporaries, ``E`` and ``D`` are written directly in the layout the MLPG kernels read.
"""
import ctypes

import numpy as np
from scipy import linalg

from ..paramgen import mlpg_batch


def _compute_precision_cholesky_full(covariances):
    """Upper factors U with U U^T = covariance^-1, as scikit-learn stores them (gmm.py:8-41)."""
    n_components, n_features, _ = covariances.shape
    out = np.empty((n_components, n_features, n_features))
    for k, cov in enumerate(covariances):
        try:
            c = linalg.cholesky(cov, lower=True)
        except linalg.LinAlgError:
            raise ValueError(
                "Fitting the mixture model failed because some components have ill-defined empirical "
                "covariance (for instance caused by singleton or collapsed samples). Try to decrease the "
                "number of components, or increase reg_covar.")
        out[k] = linalg.solve_triangular(c, np.eye(n_features), lower=True).T
    return out


class MLPGBase(object):
    """Frame-wise GMM mapping ``E[p(y | x)]`` (baseline/gmm.py:47-121)."""

    def __init__(self, gmm, swap=False, diff=False):
        assert gmm.covariance_type == "full"
        D = gmm.means_.shape[1] // 2  # static + delta dim
        self.num_mixtures = gmm.means_.shape[0]
        self.weights = gmm.weights_
        self.src_means = gmm.means_[:, :D]
        self.tgt_means = gmm.means_[:, D:]
        self.covarXX = gmm.covariances_[:, :D, :D]
        self.covarXY = gmm.covariances_[:, :D, D:]
        self.covarYX = gmm.covariances_[:, D:, :D]
        self.covarYY = gmm.covariances_[:, D:, D:]
        if diff:  # GMM -> DIFFGMM (gmm.py:63-67)
            self.tgt_means = self.tgt_means - self.src_means
            self.covarYY = self.covarXX + self.covarYY - self.covarXY - self.covarYX
            self.covarXY = self.covarXY - self.covarXX
            self.covarYX = self.covarXY.transpose(0, 2, 1)
        if swap:  # (gmm.py:70-73)
            self.tgt_means, self.src_means = self.src_means, self.tgt_means
            self.covarYY, self.covarXX = self.covarXX, self.covarYY
            self.covarYX, self.covarXY = self.covarXY, self.covarYX
        self._prec_chol = _compute_precision_cholesky_full(self.covarXX)
        self._dev = None

    # ---- device-side constants -------------------------------------------------------------------
    def _constants(self):
        import torch

        from .. import _device as dev
        dev.require_cuda()
        device = torch.device("cuda", torch.cuda.current_device())
        if self._dev is not None and self._dev["device"] == device:
            return self._dev

        def t(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
        A = np.stack([np.linalg.solve(self.covarXX[m].T, self.covarYX[m].T).T for m in range(self.num_mixtures)])
        dim = self.src_means.shape[1]
        log_det = np.sum(np.log(np.diagonal(self._prec_chol, axis1=1, axis2=2)), axis=1)
        # Eq. (23) with diagonal covariances (gmm.py:239-244)
        Dm = np.stack([np.diag(self.covarYY[m]) - np.diag(self.covarYX[m]) / np.diag(self.covarXX[m]) * np.diag(self.covarXY[m])
                       for m in range(self.num_mixtures)])
        from .. import _lib
        tabs = {
            "src_means": t(self.src_means), "tgt_means": t(self.tgt_means), "prec_chol": t(self._prec_chol),
            "log_const": t(log_det + np.log(self.weights) - 0.5 * dim * np.log(2.0 * np.pi)),
            "A_t": t(A.transpose(0, 2, 1)), "Dm": t(Dm),
        }
        g = _lib.NnkGmm()
        for k, v in tabs.items():
            setattr(g, k, v.data_ptr())
        g.M, g.D = self.num_mixtures, dim
        self._dev = {"device": device, "tabs": tabs, "gmm": g}
        return self._dev

    def _to_device(self, src):
        import torch
        c = self._constants()
        return torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64)).to(c["device"]), c

    def _weighted_log_prob(self, x, c):
        """log w_m + log N(x_t | mu_m, Sigma_xx,m) for every frame and mixture, (T, M) (Eq. 9 before
        normalisation; the reference: sklearn predict_proba per frame, gmm.py:116-118)."""
        import torch

        from .. import _device as dev
        from .. import _lib
        T = x.shape[0]
        lp = torch.empty((T, self.num_mixtures), dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib.nnk_gmm_logprob(ctypes.byref(c["gmm"]), x.data_ptr(), x.stride(0), T, lp.data_ptr(),
                                            dev.current_stream_ptr(x.device)), "nnk_gmm_logprob")
        return lp

    def _map(self, x, c, mode, want_var=False):
        """mode 0: (E, D, mix) of the arg-max mixture sequence (Eq. 37, 22, 23); mode 1: posterior mean (Eq. 13)."""
        import torch

        from .. import _device as dev
        from .. import _lib
        T, D = x.shape
        lp = self._weighted_log_prob(x, c)
        E = torch.empty((T, D), dtype=torch.float64, device=x.device)
        Dv = torch.empty((T, D), dtype=torch.float64, device=x.device) if want_var else None
        _lib.check(_lib.lib.nnk_gmm_map(ctypes.byref(c["gmm"]), x.data_ptr(), x.stride(0), T, lp.data_ptr(), mode, E.data_ptr(),
                                        Dv.data_ptr() if want_var else None, None, dev.current_stream_ptr(x.device)),
                   "nnk_gmm_map")
        return E, Dv

    def transform(self, src):
        src = np.asarray(src)
        if src.ndim == 2:
            tgt = np.zeros_like(src)
            if len(src):
                tgt[:] = self._transform_frames(src)  # zeros_like keeps the dtype of src (gmm.py:89-93)
            return tgt
        return self._transform_frames(src[None])[0]

    def _transform_frame(self, src):
        """``E[p(y | x)]`` of one frame (gmm.py:97-121)."""
        return self._transform_frames(np.asarray(src)[None])[0]

    def _transform_frames(self, src):
        xa, c = self._to_device(src)
        if xa.shape[0] == 0:
            return xa.cpu().numpy()
        E, _ = self._map(xa.contiguous(), c, 1)  # Eq. (9), (11), (13) in one pass per frame tile
        return E.cpu().numpy()


class MLPG(MLPGBase):
    """Maximum likelihood parameter generation for GMM-based voice conversion (baseline/gmm.py:124-247).

    Args:
        gmm (sklearn.mixture.GaussianMixture): joint GMM of source and target features.
        windows (list): window triples, see :func:`nnmnkwii_b200.paramgen.mlpg`.
        swap (bool): if True source -> target, otherwise target -> source.
        diff (bool): convert GMM -> DIFFGMM if True.
    """

    def __init__(self, gmm, windows=None, swap=False, diff=False):
        super(MLPG, self).__init__(gmm, swap, diff)
        if windows is None:
            windows = [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5]))]
        self.windows = windows
        self.static_dim = gmm.means_.shape[-1] // 2 // len(windows)

    def _means_vars(self, x, c):
        """E (Eq. 22) and D (Eq. 23) of the sub-optimum mixture sequence (Eq. 37), on the device."""
        return self._map(x.contiguous(), c, 0, want_var=True)

    def transform(self, src):
        """Source feature sequence ``(T, D)`` -> converted static features ``(T, static_dim)``."""
        src = np.asarray(src)
        T, feature_dim = src.shape[0], src.shape[1]
        if feature_dim == self.static_dim:
            return super(MLPG, self).transform(src)
        x, c = self._to_device(src)
        E, Dv = self._means_vars(x, c)
        return mlpg_batch(E, Dv, self.windows, lengths=[T]).cpu().numpy()

    def transform_batch(self, srcs):
        """Additive: convert a list of utterances in one pass (one posterior / mapping evaluation and
        ONE batched MLPG launch for all of them).  Returns a list of ``(T_i, static_dim)`` arrays."""
        import torch
        lens = [len(s) for s in srcs]
        if not lens:
            return []
        flat = np.concatenate([np.asarray(s) for s in srcs], axis=0)
        if flat.shape[1] == self.static_dim:
            y = MLPGBase._transform_frames(self, flat)
        else:
            x, c = self._to_device(flat)
            E, Dv = self._means_vars(x, c)
            y = mlpg_batch(E, Dv, self.windows, lengths=lens).cpu().numpy()
        del torch
        off = np.concatenate([[0], np.cumsum(lens)])
        return [y[off[i]:off[i + 1]] for i in range(len(lens))]


__all__ = ["MLPGBase", "MLPG"]
