"""Device-side plumbing shared by the Python host layer: PyTorch owns allocation and streams,
libnnk_b200 (C ABI) does the arithmetic.  No numerical work happens in this file."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import CHAIN_DTYPE, NnkMlpgArgs, NnkStatus, lib

_ws_cache = {}
_chain_cache = {}

WORKSPACE_CAP_BYTES = 2 << 30  # the launcher splits a batch into waves if it needs more


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "nnmnkwii_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback. "
            "torch.cuda.is_available() is False.")


def current_stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(device, nbytes):
    """Grow-only per-device scratch (uint8 tensor)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _ws_cache.pop(key, None)
        buf = torch.empty(int(nbytes + nbytes // 8 + (1 << 20)), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def simple_chains(static_dim):
    """Chain table of a single stream laid out like the reference: window w of static dim d lives
    in column w * static_dim + d (paramgen/_mlpg.py:187)."""
    ch = np.zeros(static_dim, dtype=CHAIN_DTYPE)
    ch["in_col"] = np.arange(static_dim)
    ch["win_stride"] = static_dim
    ch["out_col"] = np.arange(static_dim)
    return ch


def chains_on_device(chains_np, device):
    key = (chains_np.tobytes(), str(device))
    t = _chain_cache.get(key)
    if t is None:
        t = torch.from_numpy(chains_np.view(np.int32).reshape(-1, 4).copy()).to(device)
        if len(_chain_cache) > 64:
            _chain_cache.clear()
        _chain_cache[key] = t
    return t


def torch_dtype_code(dt):
    if dt == torch.float32:
        return _lib.NNK_F32
    if dt == torch.float64:
        return _lib.NNK_F64
    raise TypeError("CUDA kernels support float32 / float64, got %s" % dt)


def raise_if_failed(status_word_tensor):
    """Synchronising check of the device status word -> numpy.linalg.LinAlgError like the reference
    (scipy.linalg.LinAlgError is the same class; _bandmat/linalg.pyx:79-82)."""
    word = int(status_word_tensor.item()) & 0xFFFFFFFFFFFFFFFF
    if word:
        st = NnkStatus()
        lib.nnk_status_decode(ctypes.c_uint64(word), ctypes.byref(st))
        raise np.linalg.LinAlgError(
            "%d-th leading minor not positive definite (utterance %d, chain %d)" % (st.frame, st.utt, st.chain))


def run_mlpg(mode, *, means, variances, rhs, out, offsets, lengths, order, chains, n_chain, max_T, windows_c,
             in_ld, var_ld, go_ld, out_ld, dtype_code, go_f64, n_utt, device, check=True):
    """Fill nnk_mlpg_args_t and enqueue nnk_mlpg_{fwd,grad,solve} on torch's current stream."""
    a = NnkMlpgArgs()
    a.means = means.data_ptr() if means is not None else None
    a.vars = variances.data_ptr()
    a.grad_out = rhs.data_ptr() if rhs is not None else None
    a.out = out.data_ptr()
    a.dtype = dtype_code
    a.n_utt = n_utt
    a.in_ld, a.var_ld, a.go_ld, a.out_ld = in_ld, var_ld, go_ld, out_ld
    a.utt_off = offsets.data_ptr()
    a.utt_len = lengths.data_ptr() if lengths is not None else None
    a.order = order.data_ptr() if order is not None else None
    a.chains = chains.data_ptr()
    a.n_chain = n_chain
    a.max_T = max_T
    a.go_f64 = go_f64
    a.win = windows_c
    need = lib.nnk_mlpg_workspace_bytes(n_utt, n_chain, max_T, ctypes.byref(windows_c))
    if need == 0 and n_utt and n_chain and max_T:
        raise NotImplementedError("window set not supported by the CUDA kernels")
    groups = (n_chain + 31) // 32
    per_utt = need // max(1, n_utt)
    nbytes = max(per_utt, min(need, max(WORKSPACE_CAP_BYTES, per_utt)))
    ws = workspace(device, max(nbytes, 256))
    a.workspace = ws.data_ptr()
    a.workspace_bytes = ws.numel()
    status = torch.zeros(1, dtype=torch.int64, device=device)
    a.status_word = status.data_ptr()
    fn = {"fwd": lib.nnk_mlpg_fwd, "grad": lib.nnk_mlpg_grad, "solve": lib.nnk_mlpg_solve}[mode]
    _lib.check(fn(ctypes.byref(a), current_stream_ptr(device)), "nnk_mlpg_" + mode)
    del groups
    if check:
        raise_if_failed(status)
    return status
