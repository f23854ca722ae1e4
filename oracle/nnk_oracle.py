"""ctypes binding of oracle/nnk_oracle.c (+ helpers to reach the real reference in oracle/_ref).

TEST INFRASTRUCTURE, NOT PRODUCT -- see the header of nnk_oracle.c.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "nnk_oracle.c")
_SO = os.path.join(_HERE, "_build", "libnnk_oracle.so")
_lib = None

__all__ = [
    "build", "lib", "mlpg", "mlpg_grad", "unit_variance_mlpg_matrix", "cost", "dtw", "fastdtw",
    "expand_window", "melcd", "delta_features", "cholesky_banded_lower", "cho_solve_lower", "cholesky_inv_banded", "trim_zeros_frames_len", "import_reference", "reference_available",
    "LOGDB_CONST",
]

LOGDB_CONST = 10.0 / np.log(10.0) * np.sqrt(2.0)  # metrics/__init__.py:5


def build(force=False):
    """gcc the C restatement into oracle/_build/ (git-ignored, travels with gpurun)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if (not force and os.path.exists(_SO)
            and (not os.path.exists(_SRC) or os.path.getmtime(_SO) >= os.path.getmtime(_SRC))):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", _SRC, "-o", _SO, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        c_long, c_int, vp = ctypes.c_long, ctypes.c_int, ctypes.c_void_p
        L.orc_mlpg.restype = c_long
        L.orc_mlpg.argtypes = [vp, vp, c_int, c_int, c_long, c_long, c_int, vp, vp, vp, vp]
        L.orc_mlpg_grad.restype = c_long
        L.orc_mlpg_grad.argtypes = [vp, c_int, c_int, c_long, c_long, c_int, vp, vp, vp, vp, c_int, vp]
        L.orc_unit_variance_mlpg_matrix.restype = c_long
        L.orc_unit_variance_mlpg_matrix.argtypes = [c_long, c_int, vp, vp, vp, vp]
        L.orc_cost.restype = ctypes.c_double
        L.orc_cost.argtypes = [vp, vp, c_long, c_int]
        L.orc_logdb_const.restype = ctypes.c_double
        L.orc_fastdtw.restype = c_long
        L.orc_fastdtw.argtypes = [vp, vp, c_long, c_long, c_long, c_int, c_long, vp, vp, vp, vp]
        L.orc_dtw_window.restype = c_long
        L.orc_dtw_window.argtypes = [vp, vp, c_long, c_long, c_long, c_int, vp, vp, vp, vp, vp, vp]
        L.orc_cholesky_banded_lower.restype = c_long
        L.orc_cholesky_banded_lower.argtypes = [vp, c_long, c_long]
        L.orc_cho_solve_lower.restype = c_long
        L.orc_cho_solve_lower.argtypes = [vp, c_long, c_long, vp, vp]
        L.orc_cholesky_inv_banded.restype = None
        L.orc_cholesky_inv_banded.argtypes = [vp, c_long, c_long, vp]
        L.orc_expand_window.restype = None
        L.orc_expand_window.argtypes = [vp, vp, c_long, c_long, c_long, c_long, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _windows(windows):
    wl = np.asarray([int(w[0]) for w in windows], dtype=np.int32)
    wu = np.asarray([int(w[1]) for w in windows], dtype=np.int32)
    for l, u, c in windows:
        assert l >= 0 and u >= 0 and len(c) == l + u + 1  # _mlpg.py:44-45
    coef = np.ascontiguousarray(np.concatenate([np.asarray(w[2], dtype=np.float64).ravel() for w in windows]))
    return wl, wu, coef


class LinAlgError(np.linalg.LinAlgError):
    pass


def _check(status):
    if status > 0:
        raise LinAlgError("%d-th leading minor not positive definite" % status)
    if status < 0:
        raise LinAlgError("singular matrix / bad arguments (status %d)" % status)


def _dt(a):
    if a.dtype == np.float32:
        return 0
    if a.dtype == np.float64:
        return 1
    raise TypeError("oracle supports float32/float64, got %s" % a.dtype)


def mlpg(mean_frames, variance_frames, windows):
    """paramgen.mlpg restated (nnk_oracle.c: orc_mlpg)."""
    mean_frames = np.ascontiguousarray(mean_frames)
    variance_frames = np.ascontiguousarray(variance_frames, dtype=mean_frames.dtype)
    T, D = mean_frames.shape
    var1d = variance_frames.ndim == 1 and variance_frames.shape[0] == D
    if not var1d:
        assert mean_frames.shape == variance_frames.shape
    wl, wu, coef = _windows(windows)
    out = np.zeros((T, D // len(windows)), dtype=mean_frames.dtype)
    _check(lib().orc_mlpg(_p(mean_frames), _p(variance_frames), _dt(mean_frames), int(var1d), T, D,
                          len(windows), _p(wl), _p(wu), _p(coef), _p(out)))
    return out


def mlpg_grad(mean_frames, variance_frames, windows, grad_output):
    """paramgen.mlpg_grad restated in closed form (nnk_oracle.c: orc_mlpg_grad)."""
    T, D = mean_frames.shape
    variance_frames = np.ascontiguousarray(variance_frames)
    var1d = variance_frames.ndim == 1
    grad_output = np.ascontiguousarray(grad_output)
    wl, wu, coef = _windows(windows)
    out = np.zeros((T, D), dtype=np.float32)
    _check(lib().orc_mlpg_grad(_p(variance_frames), _dt(variance_frames), int(var1d), T, D, len(windows),
                               _p(wl), _p(wu), _p(coef), _p(grad_output), _dt(grad_output), _p(out)))
    return out


def unit_variance_mlpg_matrix(windows, T):
    wl, wu, coef = _windows(windows)
    R = np.zeros((T, len(windows) * T), dtype=np.float32)
    _check(lib().orc_unit_variance_mlpg_matrix(T, len(windows), _p(wl), _p(wu), _p(coef), _p(R)))
    return R


_KIND = {"euclid": 0, "melcd": 1}


def cost(x, y, kind="melcd"):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().orc_cost(_p(x), _p(y), x.shape[0], _KIND[kind])


def fastdtw(x, y, radius=1, kind="euclid"):
    """Returns (distance, path_i, path_j, cells).  radius < 0 => exact DTW."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if x.ndim == 1:
        x, y = x[:, None], y[:, None]
    Tx, D = x.shape
    Ty = y.shape[0]
    pi = np.zeros(Tx + Ty, dtype=np.int32)
    pj = np.zeros(Tx + Ty, dtype=np.int32)
    dist = ctypes.c_double(0.0)
    cells = ctypes.c_int64(0)
    n = lib().orc_fastdtw(_p(x), _p(y), Tx, Ty, D, _KIND[kind], radius, _p(pi), _p(pj),
                          ctypes.byref(dist), ctypes.byref(cells))
    if n < 0:
        raise RuntimeError("orc_fastdtw failed: %d" % n)
    return dist.value, pi[:n].copy(), pj[:n].copy(), cells.value


def dtw(x, y, kind="euclid"):
    return fastdtw(x, y, radius=-1, kind=kind)


def expand_window(path_i, path_j, len_x, len_y, radius):
    pi = np.ascontiguousarray(path_i, dtype=np.int32)
    pj = np.ascontiguousarray(path_j, dtype=np.int32)
    lo = np.zeros(len_x, dtype=np.int32)
    hi = np.zeros(len_x, dtype=np.int32)
    lib().orc_expand_window(_p(pi), _p(pj), len(pi), len_x, len_y, radius, _p(lo), _p(hi))
    return lo, hi


def cholesky_banded_lower(half_band):
    """_bandmat.linalg._cholesky_banded(mat, lower=True) (linalg.pyx:36-104)."""
    a = np.array(half_band, dtype=np.float64, order="C")
    _check(lib().orc_cholesky_banded_lower(_p(a), a.shape[0] - 1, a.shape[1]))
    return a


def cho_solve_lower(chol, b):
    chol = np.ascontiguousarray(chol, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    _check(lib().orc_cho_solve_lower(_p(chol), chol.shape[0] - 1, chol.shape[1], _p(b), _p(x)))
    return x


def cholesky_inv_banded(L_full, width):
    """util.linalg.cholesky_inv_banded (util/_linalg.pyx:45-71)."""
    L_full = np.ascontiguousarray(L_full, dtype=np.float64)
    T = L_full.shape[0]
    P = np.zeros((T, T))
    lib().orc_cholesky_inv_banded(_p(L_full), T, width, _p(P))
    return P


def melcd(X, Y, lengths=None):
    """metrics.melcd restated for numpy inputs (metrics/__init__.py:52-71)."""
    X, Y = np.asarray(X), np.asarray(Y)
    if lengths is None:
        z = X - Y
        r = np.sqrt((z * z).sum(-1))
        if not np.isscalar(r):
            r = r.mean()
        return LOGDB_CONST * float(r)
    if X.ndim == 2:
        X, Y = X[:, :, None], Y[:, :, None]
    s, T = 0.0, np.sum(lengths)
    for x, y, length in zip(X, Y, lengths):
        z = x[:length] - y[:length]
        s += np.sqrt((z * z).sum(-1)).sum()
    return LOGDB_CONST * float(s) / float(T)


def delta_features(x, windows):
    """preprocessing.delta_features restated (nnmnkwii/preprocessing/generic.py:229-288):
    per window and static dim ``np.correlate(x[:, d], window, mode="same")`` into a result of x's dtype."""
    T, D = x.shape
    assert len(windows) > 0
    out = np.empty((T, D * len(windows)), dtype=x.dtype)
    for idx, w in enumerate(windows):
        window = w[2] if isinstance(w, tuple) else w
        y = np.zeros_like(x)
        for d in range(D):
            y[:, d] = np.correlate(x[:, d], window, mode="same")
        out[:, D * idx:D * idx + D] = y
    return out


def trim_zeros_frames_len(x, eps=1e-7):
    """Length kept by preprocessing.trim_zeros_frames(x, eps, trim='b') (generic.py:312-323)."""
    s = np.sum(np.abs(x), axis=1)
    s[s < eps] = 0.0
    return len(np.trim_zeros(s, trim="b"))


def reference_available():
    return os.path.isdir(os.path.join(_HERE, "_ref", "nnmnkwii"))


def import_reference():
    """Import the UNMODIFIED reference from oracle/_ref (built by oracle/build_ref.sh)."""
    ref = os.path.join(_HERE, "_ref")
    if not reference_available():
        raise ImportError("oracle/_ref not built: run oracle/build_ref.sh where /root/reference exists")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import nnmnkwii  # noqa: F401
        import nnmnkwii.paramgen  # noqa: F401
    return nnmnkwii
