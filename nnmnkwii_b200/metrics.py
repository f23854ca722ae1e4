"""``melcd`` -- drop-in for ``nnmnkwii.metrics.melcd`` (nnmnkwii/metrics/__init__.py:27-71).

As a *metric* over arrays it is a handful of elementwise operations; it is written here with the
array's own operators, so NumPy inputs are evaluated by NumPy and torch (CUDA) tensors by torch on
their device, exactly as in the reference.  As the DTW *local cost* (``DTWAligner(dist=melcd)``) it is
never called per cell: the aligner recognises it and the sm_100a wavefront kernel evaluates
``10/ln10 * sqrt(2) * ||x - y||_2`` in registers (csrc/nnk_dtw.cu).
"""
import math

import numpy as np

_logdb_const = 10.0 / np.log(10.0) * np.sqrt(2.0)  # metrics/__init__.py:5


def _sqrt(x):
    isnumpy = isinstance(x, np.ndarray)
    isscalar = np.isscalar(x)
    return np.sqrt(x) if isnumpy else math.sqrt(x) if isscalar else x.sqrt()


def _sum(x):
    if isinstance(x, list) or isinstance(x, np.ndarray):
        return np.sum(x)
    return float(x.sum())


def melcd(X, Y, lengths=None):
    """Mel-cepstrum distortion (MCD) in dB.

    Args:
        X, Y: shape ``(D,)``, ``(T, D)`` or ``(B, T, D)``; NumPy arrays or torch tensors.
        lengths (list): lengths of padded inputs (mini-batch case).

    Returns:
        float: mean mel-cepstrum distortion in dB.
    """
    if lengths is None:
        z = X - Y
        r = _sqrt((z * z).sum(-1))
        if not np.isscalar(r):
            r = r.mean()
        return _logdb_const * float(r)

    if len(X.shape) == 2:
        X, Y = X[:, :, None], Y[:, :, None]

    s = 0.0
    T = _sum(lengths)
    for x, y, length in zip(X, Y, lengths):
        x, y = x[:length], y[:length]
        z = x - y
        s += _sqrt((z * z).sum(-1)).sum()

    return _logdb_const * float(s) / float(T)


__all__ = ["melcd"]
