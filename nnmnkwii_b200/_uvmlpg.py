"""Host side of UnitVarianceMLPG: reduce the dense MLPG matrix R to its numerical band (once per R,
cached) and enqueue the stencil sweeps of csrc/nnk_uvmlpg.cu.  No arithmetic on the host."""
import ctypes
import weakref

import torch

from . import _device as dev
from . import _lib
from ._lib import lib

# relative magnitude (w.r.t. max |R|) below which off-band entries of R are dropped.  R is float32
# (resolution 6e-8); everything dropped sums to well below that.  K = T - 1 reproduces R exactly.
BAND_REL_TOL = 2.0 ** -32

_band_cache = {}


# rows whose band differs from the middle row by less than this (relative to max |R|) share its filter
TOEPLITZ_REL_TOL = 2.0 ** -22
TOEPLITZ_MIN_ROWS = 64


class Band(object):
    __slots__ = ("Rb", "RbT", "K", "T", "nw", "dtype", "toep", "toepT")


def _toeplitz_interval(table, peak):
    """(t_lo, t_hi, taps) of the maximal run of rows around T//2 equal to the middle row (host side,
    once per R): table is the (T, nw, 2K+1) band table."""
    import numpy as np
    tab = table.cpu().numpy()
    T = tab.shape[0]
    mid = T // 2
    ok = np.abs(tab - tab[mid]).reshape(T, -1).max(axis=1) <= peak * TOEPLITZ_REL_TOL
    lo = mid
    while lo > 0 and ok[lo - 1]:
        lo -= 1
    hi = mid + 1
    while hi < T and ok[hi]:
        hi += 1
    return lo, hi, np.ascontiguousarray(tab[mid], dtype=np.float32)


def band_of(R, device):
    """Band tables of R (cached on (data_ptr, version, shape, device))."""
    assert R.dim() == 2 and R.shape[1] % R.shape[0] == 0
    key = (R.data_ptr(), R._version, tuple(R.shape), R.dtype, str(R.device), str(device))
    hit = _band_cache.get(key)
    if hit is not None and hit[0]() is R:
        return hit[1]
    Rd = R.detach().to(device)
    if Rd.dtype not in (torch.float32, torch.float64):
        Rd = Rd.to(torch.float32)
    Rd = Rd.contiguous()
    T = Rd.shape[0]
    nw = Rd.shape[1] // T
    code = dev.torch_dtype_code(Rd.dtype)
    stream = dev.current_stream_ptr(device)
    profile = torch.empty(T, dtype=torch.float32, device=device)
    _lib.check(lib.nnk_uv_band_profile(Rd.data_ptr(), code, T, nw, profile.data_ptr(), stream), "nnk_uv_band_profile")
    prof = profile.cpu()
    peak = float(prof.max())
    above = torch.nonzero(prof > peak * BAND_REL_TOL)
    K = int(above.max()) if above.numel() else 0
    b = Band()
    b.K, b.T, b.nw, b.dtype = K, T, nw, Rd.dtype
    b.Rb = torch.empty((T, nw, 2 * K + 1), dtype=Rd.dtype, device=device)
    b.RbT = torch.empty((T, nw, 2 * K + 1), dtype=Rd.dtype, device=device)
    _lib.check(lib.nnk_uv_band_extract(Rd.data_ptr(), code, T, nw, K, b.Rb.data_ptr(), b.RbT.data_ptr(), stream),
               "nnk_uv_band_extract")
    b.toep = b.toepT = None
    if Rd.dtype == torch.float32 and nw <= 3 and K <= 64 and T >= TOEPLITZ_MIN_ROWS:
        lo, hi, taps = _toeplitz_interval(b.Rb, peak)
        if hi - lo >= TOEPLITZ_MIN_ROWS:
            b.toep = (lo, hi, taps)
        lo, hi, taps = _toeplitz_interval(b.RbT, peak)
        if hi - lo >= TOEPLITZ_MIN_ROWS:
            b.toepT = (lo, hi, taps)
    if len(_band_cache) > 16:
        _band_cache.clear()
    try:
        _band_cache[key] = (weakref.ref(R), b)
    except TypeError:
        pass
    return b


def apply_forward(band, means3, reshaped):
    """means3: (B, T, nw*sd) [reshaped=False] or (B, nw*T, sd) [reshaped=True] -> (B, T, sd)."""
    B = means3.shape[0]
    T, nw = band.T, band.nw
    if reshaped:
        assert means3.shape[1] == nw * T
        sd = means3.shape[2]
    else:
        assert means3.shape[1] == T
        sd = means3.shape[2] // nw
        if means3.shape[2] != nw * sd:
            means3 = means3[..., : nw * sd]
    x = means3.to(band.dtype).contiguous()
    y = torch.empty((B, T, sd), dtype=band.dtype, device=x.device)
    if band.toep is not None:
        lo, hi, taps = band.toep
        _lib.check(lib.nnk_uv_apply_toeplitz(band.Rb.data_ptr(), taps.ctypes.data, x.data_ptr(), y.data_ptr(), B, T, sd, nw,
                                             band.K, lo, hi, 0, int(reshaped), dev.current_stream_ptr(x.device)),
                   "nnk_uv_apply_toeplitz")
    else:
        _lib.check(lib.nnk_uv_apply(band.Rb.data_ptr(), x.data_ptr(), y.data_ptr(), dev.torch_dtype_code(band.dtype),
                                    B, T, sd, nw, band.K, 0, int(reshaped), dev.current_stream_ptr(x.device)), "nnk_uv_apply")
    return y.to(means3.dtype) if y.dtype != means3.dtype else y


def apply_backward(band, grad_output3, reshaped, D):
    """grad_output3: (B, T, sd) -> gradient w.r.t. means: (B, nw*T, sd) or (B, T, D)."""
    B, T, sd = grad_output3.shape
    nw = band.nw
    assert T == band.T
    g = grad_output3.to(band.dtype).contiguous()
    if reshaped:
        gx = torch.empty((B, nw * T, sd), dtype=band.dtype, device=g.device)
    else:
        gx = torch.empty((B, T, nw * sd), dtype=band.dtype, device=g.device)
    if band.toepT is not None:
        lo, hi, taps = band.toepT
        _lib.check(lib.nnk_uv_apply_toeplitz(band.RbT.data_ptr(), taps.ctypes.data, g.data_ptr(), gx.data_ptr(), B, T, sd, nw,
                                             band.K, lo, hi, 1, int(reshaped), dev.current_stream_ptr(g.device)),
                   "nnk_uv_apply_toeplitz")
    else:
        _lib.check(lib.nnk_uv_apply(band.RbT.data_ptr(), g.data_ptr(), gx.data_ptr(), dev.torch_dtype_code(band.dtype),
                                    B, T, sd, nw, band.K, 1, int(reshaped), dev.current_stream_ptr(g.device)), "nnk_uv_apply")
    if not reshaped and nw * sd != D:  # trailing columns the forward ignored get zero gradient
        full = torch.zeros((B, T, D), dtype=band.dtype, device=g.device)
        full[..., : nw * sd] = gx
        gx = full
    return gx.to(grad_output3.dtype) if gx.dtype != grad_output3.dtype else gx


_ = ctypes
