"""Parameter generation (MLPG) -- drop-in for ``nnmnkwii.paramgen`` on a B200.

Same names, argument order, defaults, return types and error behaviour as the reference
(nnmnkwii/paramgen/_mlpg.py, ``__init__.py:1-17``):

    build_win_mats(windows, T)                          _mlpg.py:13-50
    mlpg(mean_frames, variance_frames, windows)         _mlpg.py:92-199
    mlpg_grad(mean_frames, variance_frames, windows, grad_output)   _mlpg.py:202-281
    full_window_mat(win_mats, T)                        _mlpg.py:284-294 (mlpg_helper.pyx:10-32)
    unit_variance_mlpg_matrix(windows, T)               _mlpg.py:297-373
    reshape_means(means, static_dim)                    _mlpg.py:376-405

All arithmetic runs in the sm_100a kernels of libnnk_b200 (csrc/nnk_mlpg.cu) through the C ABI in
include/nnk_b200.h.  NumPy inputs go through the host-buffer entry points (copies included); torch
CUDA tensors are used in place on the current stream and a CUDA tensor comes back.  There is no CPU
fallback: without the library / a GPU these functions raise.

Additive, batched entry points (the reference has none -- its notebooks loop over utterances and
streams in Python): :class:`StreamLayout`, :func:`merlin_layout`, :func:`mlpg_batch`.
"""
import ctypes

import numpy as np

from . import _lib
from .bandmat import BandMat

__all__ = [
    "mlpg_grad_batch",
    "build_win_mats", "mlpg", "mlpg_grad", "full_window_mat", "unit_variance_mlpg_matrix", "reshape_means",
    "StreamLayout", "merlin_layout", "mlpg_batch",
]


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def build_win_mats(windows, T):
    """Builds a window matrix of a given size for each window in a collection (_mlpg.py:13-50).

    Returns a list of ``T x T`` Toeplitz :class:`~nnmnkwii_b200.bandmat.BandMat` (lower bandwidth
    ``l``, upper bandwidth ``u``, ``transposed=True`` exactly like the reference's
    ``bm.band_c_bm(u, l, win_coeffs).T``).  Host-side only: the kernels take the window
    coefficients directly and never build these matrices.
    """
    win_mats = []
    for ll, u, win_coeff in windows:
        assert ll >= 0 and u >= 0
        assert len(win_coeff) == ll + u + 1
        win_coeffs = np.tile(np.reshape(win_coeff, (ll + u + 1, 1)), T)
        win_mats.append(BandMat(u, ll, win_coeffs.copy()).T)
    return win_mats


def full_window_mat(win_mats, T):
    """Dense ``(T * num_windows, T)`` float64 stack of the window matrices (_mlpg.py:284-294)."""
    mat_full = np.zeros((T * len(win_mats), T))
    for win_index, win_mat in enumerate(win_mats):
        mat_full[win_index * T:(win_index + 1) * T, :] = win_mat.full()
    return mat_full


def reshape_means(means, static_dim):
    """Reshape means (``T x D``) to (``T*num_windows x static_dim``); no-op if already reshaped
    (_mlpg.py:376-405)."""
    T, D = means.shape
    if D == static_dim:
        return means
    if _is_torch(means):
        return means.reshape(T, -1, static_dim).transpose(0, 1).reshape(-1, static_dim)
    return means.reshape(T, -1, static_dim).transpose(1, 0, 2).reshape(-1, static_dim)


# ---------------------------------------------------------------------------------------------------
# layouts (additive)
# ---------------------------------------------------------------------------------------------------
class StreamLayout(object):
    """Where the streams of a ``(T, D)`` frame matrix live.

    ``streams`` is a list of ``(in_col, static_dim)`` for smoothed streams (window ``w`` of static
    dimension ``d`` is column ``in_col + w * static_dim + d``, as in the reference) or
    ``(in_col, static_dim, "copy")`` for columns that are passed through (e.g. Merlin's vuv flag).
    Output columns are assigned consecutively in the order given.
    """

    def __init__(self, D_in, streams):
        self.D_in = int(D_in)
        rows = []
        out_col = 0
        self.slices = []
        for s in streams:
            in_col, sd = int(s[0]), int(s[1])
            copy = len(s) > 2 and s[2] == "copy"
            for d in range(sd):
                rows.append((in_col + d, 0 if copy else sd, out_col + d, 1 if copy else 0))
            self.slices.append((out_col, out_col + sd))
            out_col += sd
        self.D_out = out_col
        self.chains = np.array(rows, dtype=_lib.CHAIN_DTYPE) if rows else np.zeros(0, dtype=_lib.CHAIN_DTYPE)
        self.n_chain = len(rows)

    @classmethod
    def single(cls, D, num_windows):
        """One stream occupying the whole matrix: ``static_dim = D // num_windows`` (_mlpg.py:172)."""
        return cls(D, [(0, D // num_windows)])


def merlin_layout():
    """The 187-column Merlin / slt_arctic acoustic layout of the gallery notebooks: mgc 180 (static
    60), lf0 3 (static 1), vuv 1 (copied), bap 3 (static 1) -> 63 output columns."""
    return StreamLayout(187, [(0, 60), (180, 1), (183, 1, "copy"), (184, 1)])


def _offsets_from(lengths=None, offsets=None, n_rows=None):
    if offsets is not None:
        off = np.asarray(offsets, dtype=np.int64)
    elif lengths is not None:
        off = np.concatenate([[0], np.cumsum(np.asarray(lengths, dtype=np.int64))])
    else:
        off = np.array([0, n_rows], dtype=np.int64)
    return np.ascontiguousarray(off)


def _np_dtype_code(dt):
    if dt == np.float32:
        return _lib.NNK_F32
    if dt == np.float64:
        return _lib.NNK_F64
    return None


def mlpg_batch(means, variances, windows, lengths=None, offsets=None, layout=None, check=True, out=None):
    """Batched MLPG over many utterances and streams in one call (additive API).

    Args:
        means: flat ``(sum_T, D)`` frame matrix holding the utterances back to back (give ``lengths``
            or ``offsets``), or a zero-padded ``(B, Tmax, D)`` batch (give ``lengths``).
            NumPy array (host path, copies included) or torch CUDA tensor (in place, current stream).
        variances: same shape as ``means`` (per-frame) or ``(D,)`` (global).
        windows: list of ``(l, u, coeff)`` triples shared by all smoothed streams.
        layout: :class:`StreamLayout`; default = one stream covering all columns.
        out: optional preallocated ``(sum_T, D_out)`` NumPy result buffer of the input dtype (flat
            host form only); pass page-locked memory to keep the device-to-host copy asynchronous.

    Returns:
        ``(sum_T, D_out)`` (or ``(B, Tmax, D_out)``) array / tensor of the input dtype.
    """
    padded = means.ndim == 3
    D = means.shape[-1]
    if layout is None:
        layout = StreamLayout.single(D, len(windows))
    assert layout.D_in == D
    if padded:
        assert lengths is not None, "padded (B, Tmax, D) input needs lengths"
        B, Tmax = means.shape[0], means.shape[1]
    if _is_torch(means):
        return _mlpg_batch_device(means, variances, windows, lengths, offsets, layout, padded, check)

    dtype = means.dtype
    code = _np_dtype_code(dtype)
    work_dtype = dtype if (code is not None and np.asarray(variances).dtype == dtype) else np.float64
    m = np.ascontiguousarray(means, dtype=work_dtype)
    v = np.ascontiguousarray(variances, dtype=work_dtype)
    var1d = v.ndim == 1
    if var1d:
        assert v.shape[0] >= D
    else:
        assert m.shape == v.shape
    wc = _lib.make_windows(windows)
    if padded:
        # zero-padded batch: run on the flat view, one "utterance" per row block
        lens = np.asarray(lengths, dtype=np.int64)
        assert len(lens) == B and lens.max(initial=0) <= Tmax
        keep = np.concatenate([np.arange(b * Tmax, b * Tmax + lens[b]) for b in range(B)]) if B else np.zeros(0, np.int64)
        flat_m = m.reshape(B * Tmax, D)[keep]
        flat_v = v if var1d else v.reshape(B * Tmax, D)[keep]
        y = mlpg_batch(flat_m, flat_v, windows, lengths=lens, layout=layout, check=check)
        out = np.zeros((B * Tmax, layout.D_out), dtype=dtype)
        out[keep] = y
        return out.reshape(B, Tmax, layout.D_out)
    n_rows = m.shape[0]
    off = _offsets_from(lengths, offsets, n_rows)
    assert off[0] == 0 and off[-1] == n_rows
    if out is not None and (out.shape != (n_rows, layout.D_out) or out.dtype != work_dtype or not out.flags.c_contiguous):
        raise ValueError("out must be a C-contiguous (%d, %d) array of dtype %s" % (n_rows, layout.D_out, work_dtype))
    if out is None:
        out = np.empty((n_rows, layout.D_out), dtype=work_dtype)
    st = _lib.NnkStatus()
    chains = np.ascontiguousarray(layout.chains)
    rc = _lib.lib.nnk_mlpg_batch_host(
        m.ctypes.data, v.ctypes.data, int(var1d), _np_dtype_code(np.dtype(work_dtype)), n_rows, D, layout.D_out,
        off.ctypes.data, len(off) - 1, chains.ctypes.data, layout.n_chain, ctypes.byref(wc), out.ctypes.data,
        ctypes.byref(st))
    _lib.check(rc, "nnk_mlpg_batch_host")
    return out if out.dtype == dtype else out.astype(dtype)


def _mlpg_batch_device(means, variances, windows, lengths, offsets, layout, padded, check):
    import torch

    from . import _device as dev

    dev.require_cuda()
    assert means.is_cuda, "torch inputs must be CUDA tensors (no CPU fallback)"
    device = means.device
    dtype = means.dtype
    if dtype not in (torch.float32, torch.float64) or variances.dtype != dtype:
        work = torch.float64
    else:
        work = dtype
    m = means.to(work).contiguous()
    v = variances.to(device=device, dtype=work)
    var1d = v.dim() == 1
    D = m.shape[-1]
    if not var1d:
        v = v.expand_as(m).contiguous() if v.shape != m.shape else v.contiguous()
    else:
        v = v.contiguous()
    if padded:
        B, Tmax = m.shape[0], m.shape[1]
        lens_np = np.asarray(lengths.cpu() if _is_torch(lengths) else lengths, dtype=np.int64)
        off_np = np.arange(B + 1, dtype=np.int64) * Tmax
        n_rows = B * Tmax
        lens_t = torch.from_numpy(lens_np.astype(np.int32)).to(device)
        max_T = int(lens_np.max(initial=0))
    else:
        n_rows = m.shape[0]
        if _is_torch(offsets):
            offsets = offsets.cpu().numpy()
        if _is_torch(lengths):
            lengths = lengths.cpu().numpy()
        off_np = _offsets_from(lengths, offsets, n_rows)
        lens_np = np.diff(off_np)
        lens_t = None
        max_T = int(lens_np.max(initial=0))
    n_utt = len(off_np) - 1
    out = torch.zeros((n_rows, layout.D_out), dtype=work, device=device)
    if n_utt and max_T and layout.n_chain:
        order = np.argsort(-lens_np, kind="stable").astype(np.int32)
        dev.run_mlpg(
            "fwd", means=m, variances=v, rhs=None, out=out,
            offsets=torch.from_numpy(off_np).to(device), lengths=lens_t,
            order=torch.from_numpy(order).to(device), chains=dev.chains_on_device(layout.chains, device),
            n_chain=layout.n_chain, max_T=max_T, windows_c=_lib.make_windows(windows),
            in_ld=D, var_ld=0 if var1d else D, go_ld=0, out_ld=layout.D_out,
            dtype_code=dev.torch_dtype_code(work), go_f64=0, n_utt=n_utt, device=device, check=check)
    if padded:
        out = out.reshape(m.shape[0], m.shape[1], layout.D_out)
    return out if work == dtype else out.to(dtype)


# ---------------------------------------------------------------------------------------------------
# reference signatures
# ---------------------------------------------------------------------------------------------------
def mlpg(mean_frames, variance_frames, windows):
    r"""Maximum Likelihood Parameter Generation, ``f: (T, D) -> (T, static_dim)`` (_mlpg.py:92-199).

    .. math:: y = (\sum_l W_l^T P_l W_l)^{-1} \sum_l W_l^T P_l \mu_l

    Args:
        mean_frames (2darray): means, static + dynamic features, ``(T, D)``.
        variance_frames (2d or 1darray): per-frame ``(T, D)`` or global ``(D,)`` variances.
        windows (list): ``(l, u, win_coeff)`` triples.

    Returns:
        Generated static features ``(T, D // len(windows))`` in the dtype of ``mean_frames``.
    """
    if _is_torch(mean_frames):
        T, D = mean_frames.shape
        if variance_frames.dim() == 1 and variance_frames.shape[0] == D:
            pass
        else:
            assert mean_frames.shape == variance_frames.shape
        return mlpg_batch(mean_frames, variance_frames, windows, lengths=[T])
    mean_frames = np.asarray(mean_frames)
    variance_frames = np.asarray(variance_frames)
    dtype = mean_frames.dtype
    T, D = mean_frames.shape
    var1d = variance_frames.ndim == 1 and variance_frames.shape[0] == D
    if not var1d:
        assert mean_frames.shape == variance_frames.shape
    code = _np_dtype_code(dtype)
    work_dtype = dtype if (code is not None and variance_frames.dtype == dtype) else np.float64
    m = np.ascontiguousarray(mean_frames, dtype=work_dtype)
    v = np.ascontiguousarray(variance_frames, dtype=work_dtype)
    wc = _lib.make_windows(windows)
    static_dim = D // len(windows)
    y = np.zeros((T, static_dim), dtype=work_dtype)
    bad = ctypes.c_int32(0)
    rc = _lib.lib.nnk_mlpg_host(m.ctypes.data, v.ctypes.data, int(var1d), _np_dtype_code(np.dtype(work_dtype)),
                                T, D, ctypes.byref(wc), y.ctypes.data, ctypes.byref(bad))
    _lib.check(rc, "nnk_mlpg_host")
    return y if y.dtype == dtype else y.astype(dtype)


def mlpg_grad(mean_frames, variance_frames, windows, grad_output, check=True):
    r"""MLPG gradient (_mlpg.py:202-281): returns ``(T, D)`` float32,

    .. math:: g_{d,l} = P_{d,l} \, W_l (\sum_l W_l^T P_{d,l} W_l)^{-1} o_d

    evaluated as one banded solve + one stencil per static dimension (the reference solves a
    dense ``T x T`` right-hand side per (dimension, window)).
    """
    import torch

    from . import _device as dev

    dev.require_cuda()
    is_t = _is_torch(mean_frames)
    device = mean_frames.device if is_t and mean_frames.is_cuda else torch.device("cuda", torch.cuda.current_device())

    def to_dev(x, dt=None):
        t = x if _is_torch(x) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=device, dtype=dt) if dt is not None else t.to(device)

    T, D = mean_frames.shape
    v = to_dev(variance_frames)
    if v.dtype not in (torch.float32, torch.float64):
        v = v.to(torch.float64)
    if v.dim() == 2 and v.shape[0] > 1 and v.stride(0) == 0:
        v = v[0]  # v.expand(T, D) of a global variance (tests/test_autograd.py:191): keep it 1-D
    var1d = v.dim() == 1
    v = v.contiguous()
    go = to_dev(grad_output)
    if go.dtype not in (torch.float32, torch.float64):
        go = go.to(torch.float32)
    go = go.contiguous()
    nw = len(windows)
    static_dim = D // nw
    out = torch.zeros((T, D), dtype=torch.float32, device=device)
    if T and static_dim:
        chains = dev.simple_chains(static_dim)
        dev.run_mlpg(
            "grad", means=None, variances=v, rhs=go, out=out,
            offsets=torch.tensor([0, T], dtype=torch.int64, device=device), lengths=None, order=None,
            chains=dev.chains_on_device(chains, device), n_chain=static_dim, max_T=T,
            windows_c=_lib.make_windows(windows), in_ld=D, var_ld=0 if var1d else D, go_ld=go.shape[1], out_ld=D,
            dtype_code=dev.torch_dtype_code(v.dtype), go_f64=int(go.dtype == torch.float64), n_utt=1,
            device=device, check=check)
    if is_t:
        return out
    return out.cpu().numpy()


def mlpg_grad_batch(variances, windows, grad_output, lengths, layout=None, check=True):
    """Batched :func:`mlpg_grad` on the device (additive API): the gradients of
    ``mlpg_batch(means, variances, windows, lengths, layout=layout)`` with respect to ``means`` for
    every utterance of the batch in ONE launch of ``nnk_mlpg_grad``.

    Args:
        variances: CUDA tensor, flat ``(sum_T, D)`` / padded ``(B, Tmax, D)`` per-frame variances,
            or ``(D,)`` global.
        grad_output: CUDA tensor ``(sum_T, D_out)`` / ``(B, Tmax, D_out)``, the gradient with respect
            to the generated trajectories.
        lengths: frames per utterance.

    Returns:
        float32 CUDA tensor shaped like the means (``(sum_T, D)`` or ``(B, Tmax, D)``); rows beyond
        an utterance's length (padded form) are zero.
    """
    import torch

    from . import _device as dev

    dev.require_cuda()
    assert _is_torch(grad_output) and grad_output.is_cuda, "device tensors only (no CPU fallback)"
    device = grad_output.device
    padded = grad_output.dim() == 3
    lens_np = np.asarray(lengths.cpu() if _is_torch(lengths) else lengths, dtype=np.int64)
    go = grad_output.detach()
    if go.dtype not in (torch.float32, torch.float64):
        go = go.to(torch.float32)
    go = go.contiguous()
    v = variances.detach().to(device)
    if v.dtype not in (torch.float32, torch.float64):
        v = v.to(torch.float64)
    var1d = v.dim() == 1
    D = v.shape[-1]
    if layout is None:
        layout = StreamLayout.single(D, len(windows))
    assert layout.D_in == D and go.shape[-1] == layout.D_out
    if padded:
        B, Tmax = go.shape[0], go.shape[1]
        off_np = np.arange(B + 1, dtype=np.int64) * Tmax
        lens_t = torch.from_numpy(lens_np.astype(np.int32)).to(device)
        n_rows = B * Tmax
        if not var1d:
            v = v.expand(B, Tmax, D)
    else:
        off_np = _offsets_from(lens_np, None, go.shape[0])
        lens_t = None
        n_rows = go.shape[0]
        if not var1d:
            v = v.expand(n_rows, D)
    v = v.contiguous()  # materialises stride-0 expanded variances
    max_T = int(lens_np.max(initial=0))
    n_utt = len(lens_np)
    out = torch.zeros((n_rows, D), dtype=torch.float32, device=device)
    if n_utt and max_T and layout.n_chain:
        # the gradient kernel indexes grad_output by chain: chain c reads column c, so route the
        # layout's output columns to chain order (identity for a single stream)
        out_cols = torch.from_numpy(layout.chains["out_col"].astype(np.int64)).to(device)
        go2 = go.reshape(n_rows, layout.D_out)
        if not np.array_equal(layout.chains["out_col"], np.arange(layout.n_chain)):
            go2 = go2.index_select(1, out_cols).contiguous()
        dev.run_mlpg(
            "grad", means=None, variances=v, rhs=go2, out=out,
            offsets=torch.from_numpy(off_np).to(device), lengths=lens_t,
            order=torch.from_numpy(np.argsort(-lens_np, kind="stable").astype(np.int32)).to(device),
            chains=dev.chains_on_device(layout.chains, device), n_chain=layout.n_chain, max_T=max_T,
            windows_c=_lib.make_windows(windows), in_ld=D, var_ld=0 if var1d else D, go_ld=layout.n_chain, out_ld=D,
            dtype_code=dev.torch_dtype_code(v.dtype), go_f64=int(go2.dtype == torch.float64), n_utt=n_utt,
            device=device, check=check)
    return out.reshape(go.shape[0], go.shape[1], D) if padded else out


def unit_variance_mlpg_matrix(windows, T):
    r"""MLPG matrix for unit-variance inputs, ``R = (W^T W)^{-1} W^T`` with the reference's edge
    rule for dynamic windows; ``(T, num_windows * T)`` float32 (_mlpg.py:297-373).

    Built on the GPU as ``num_windows * T`` banded solves (one chain per column of
    :math:`\tilde W^T`) instead of the reference's dense ``O(T^2)`` banded inverse followed by a
    dense ``(T x T)(T x 3T)`` product.
    """
    import torch

    from . import _device as dev

    dev.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device())
    nw = len(windows)
    win_mats = build_win_mats(windows, T)
    max_win_width = int(np.max([max(w.l, w.u) for w in win_mats]))
    # right-hand sides = columns of Wtilde^T: row r of window w, edge rows of dynamic windows zeroed
    mask = np.zeros(T)
    if max_win_width > 0:  # precisions.data[:, m:-m] += 1.0 (_mlpg.py:354); m == 0 -> empty slice
        mask[max_win_width:T - max_win_width] = 1.0
    rhs = np.zeros((T, nw * T))
    for w, (l, u, c) in enumerate(windows):
        c = np.asarray(c, dtype=np.float64)
        rows = np.arange(T)
        scale = np.ones(T) if w == 0 else mask
        for k in range(-l, u + 1):
            cols = rows + k
            ok = (cols >= 0) & (cols < T)
            rhs[cols[ok], w * T + rows[ok]] = scale[ok] * c[l + k]
    n_chain = nw * T
    chains = np.zeros(n_chain, dtype=_lib.CHAIN_DTYPE)  # in_col = win_stride = 0: all read variance[0] == 1
    chains["out_col"] = np.arange(n_chain)
    out = torch.zeros((T, n_chain), dtype=torch.float64, device=device)
    if T:
        dev.run_mlpg(
            "solve", means=None, variances=torch.ones(max(1, nw), dtype=torch.float64, device=device),
            rhs=torch.from_numpy(rhs).to(device), out=out,
            offsets=torch.tensor([0, T], dtype=torch.int64, device=device), lengths=None, order=None,
            chains=dev.chains_on_device(chains, device), n_chain=n_chain, max_T=T,
            windows_c=_lib.make_windows(windows), in_ld=1, var_ld=0, go_ld=n_chain, out_ld=n_chain,
            dtype_code=_lib.NNK_F64, go_f64=1, n_utt=1, device=device, check=True)
    return out.to(torch.float32).cpu().numpy()
