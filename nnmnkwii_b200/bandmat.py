"""Minimal host-side banded-matrix container.

Only what ``paramgen.build_win_mats`` / ``paramgen.full_window_mat`` need to keep the reference's
return types (reference: nnmnkwii/paramgen/_bandmat/core.pyx:18-120, a vendored copy of
MattShannon/bandmat).  The reference's banded *arithmetic* (dot_mv/dot_mm/cholesky/solve,
tensor.pyx / linalg.pyx) is NOT re-exposed here: it lives, fused, inside the sm_100a kernels
(csrc/nnk_mlpg.cu) and never materialises a banded matrix in memory.
"""
import numpy as np


class BandMat(object):
    """A banded ``size x size`` matrix with lower bandwidth ``l`` and upper bandwidth ``u``.

    ``data`` has shape ``(l + u + 1, size)``; for a non-transposed BandMat
    ``full[i, j] = data[u + i - j, j]``; ``transposed=True`` represents the transpose of
    ``BandMat(u, l, data)`` (same conventions as core.pyx:18-67).
    """

    def __init__(self, l, u, data, transposed=False):
        self.l = int(l)
        self.u = int(u)
        self.data = data
        self.transposed = bool(transposed)
        assert self.l >= 0 and self.u >= 0
        assert self.data.ndim == 2 and self.data.shape[0] == self.l + self.u + 1

    def __repr__(self):
        return "BandMat(%r, %r, %r, transposed=%r)" % (self.l, self.u, self.data, self.transposed)

    @property
    def size(self):
        return self.data.shape[1]

    @property
    def T(self):
        # cheap: flips the flag, shares the data (core.pyx:69-78)
        return BandMat(self.u, self.l, self.data, transposed=not self.transposed)

    def full(self):
        """Dense ``(size, size)`` array."""
        n = self.size
        ll, uu = (self.u, self.l) if self.transposed else (self.l, self.u)
        out = np.zeros((n, n), dtype=self.data.dtype)
        for o in range(-uu, ll + 1):  # o = i - j
            row = uu + o
            js = np.arange(max(0, -o), max(0, n + min(0, -o)))
            out[js + o, js] = self.data[row, js]
        return out.T if self.transposed else out
