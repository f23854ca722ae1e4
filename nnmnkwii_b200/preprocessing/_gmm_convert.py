"""Frame-wise GMM mapping used by IterativeDTWAligner's conversion step.

The reference builds ``nnmnkwii.baseline.gmm.MLPG(gmm, windows=[static])`` and calls ``.transform``;
with a static-only window that class takes the frame-wise shortcut ``MLPGBase.transform``
(baseline/gmm.py:221-222, 87-121) and never reaches ``paramgen.mlpg``:

    E[m]      = mu_y[m] + Sigma_yx[m] Sigma_xx[m]^-1 (x - mu_x[m])          (Eq. 11)
    p(m | x)  = posterior of the marginal source GMM                         (Eq. 9)
    y         = sum_m p(m | x) E[m]                                          (Eq. 13)

This is host-side control logic around scikit-learn's GaussianMixture (SURVEY.md section 8a keeps the
GMM refit on the host); it is vectorised over frames but is the same arithmetic.
"""
import numpy as np
from scipy import linalg
from sklearn.mixture import GaussianMixture


class FramewiseGMMConverter(object):
    def __init__(self, gmm):
        assert gmm.covariance_type == "full"
        D = gmm.means_.shape[1] // 2
        self.num_mixtures = gmm.means_.shape[0]
        self.src_means = gmm.means_[:, :D]
        self.tgt_means = gmm.means_[:, D:]
        self.covarXX = gmm.covariances_[:, :D, :D]
        self.covarYX = gmm.covariances_[:, D:, :D]
        # marginal p(x) for the posteriors (baseline/gmm.py:75-85)
        self.px = GaussianMixture(n_components=self.num_mixtures, covariance_type="full")
        self.px.means_ = self.src_means
        self.px.covariances_ = self.covarXX
        self.px.weights_ = gmm.weights_
        chol = np.empty_like(self.covarXX)
        for k, cov in enumerate(self.covarXX):
            c = linalg.cholesky(cov, lower=True)
            chol[k] = linalg.solve_triangular(c, np.eye(D), lower=True).T
        self.px.precisions_cholesky_ = chol
        # A[m] = Sigma_yx[m] Sigma_xx[m]^-1
        self.A = np.stack([np.linalg.solve(self.covarXX[m].T, self.covarYX[m].T).T for m in range(self.num_mixtures)])

    def transform(self, src):
        src = np.atleast_2d(src)
        post = self.px.predict_proba(src)  # (T, M)
        diff = src[:, None, :] - self.src_means[None]  # (T, M, D)
        E = self.tgt_means[None] + np.einsum("mij,tmj->tmi", self.A, diff)
        return np.einsum("tm,tmi->ti", post, E)
