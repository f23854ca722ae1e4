"""Autograd functions -- drop-in for the MLPG part of ``nnmnkwii.autograd``
(nnmnkwii/autograd/_impl/mlpg.py): ``MLPG``, ``UnitVarianceMLPG``, ``mlpg``, ``unit_variance_mlpg``.

Differences that are additions, not signature changes:
  * ``MLPG`` works on CUDA tensors in place (the reference "cannot run on CUDA", mlpg.py:33) and
    moves CPU tensors to the GPU and back; its backward is one banded solve + stencil per static
    dimension instead of the reference's dense ``T x T`` solve per (dimension, window).
  * ``UnitVarianceMLPG`` keeps the ``(means, R)`` signature; the dense ``R`` (``T x nw*T``) is
    reduced once to its numerical band and applied as a banded stencil sweep
    (csrc/nnk_uvmlpg.cu) instead of two dense GEMMs that multiply ~95 % zeros.
"""
import numpy as np
import torch
from torch.autograd import Function

from . import paramgen as G


def _cuda_device(t):
    from . import _device as dev

    dev.require_cuda()
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def _global_variance(variances):
    """A global ``(D,)`` variance -- given as such or as ``v.expand(T, D)`` (stride 0 over frames, what
    the reference's ``autograd.mlpg`` builds, mlpg.py:196-197) -- as a 1-D tensor; anything else unchanged."""
    if variances.dim() == 2 and variances.shape[0] > 1 and variances.stride(0) == 0:
        return variances[0]
    return variances


class MLPG(Function):
    """Generic MLPG as an autograd function, ``f : (T, D) -> (T, static_dim)`` (mlpg.py:8-67).

    Forward = :func:`nnmnkwii_b200.paramgen.mlpg`, backward = :func:`nnmnkwii_b200.paramgen.mlpg_grad`;
    returns float32 like the reference (mlpg.py:53); gradients for ``variances`` / ``windows`` are None.
    """

    @staticmethod
    def forward(ctx, means, variances, windows):
        assert means.dim() == 2  # we cannot do MLPG on minibatch (mlpg.py:44)
        ctx.windows = windows
        variances = _global_variance(variances)  # (D,) or a stride-0 expansion of it -> 1-D (16*sd B/frame path)
        ctx.save_for_backward(means, variances)
        assert variances.dim() == 1 or means.size() == variances.size()
        device = _cuda_device(means)
        m = means.detach().to(device)
        v = variances.detach().to(device)
        # CUDA inputs: non-blocking status check (surfaces at the next call); CPU inputs (the reference's
        # use) synchronise on the copy back anyway, so the check is immediate
        ctx.check = "deferred" if means.is_cuda else True
        y = G.mlpg_batch(m, v, windows, lengths=[m.shape[0]], check=ctx.check).to(torch.float32)
        return y.to(means.device)

    @staticmethod
    def backward(ctx, grad_output):
        means, variances = ctx.saved_tensors
        device = _cuda_device(means)
        g = G.mlpg_grad(means.detach().to(device), variances.detach().to(device), ctx.windows,
                        grad_output.detach().to(device), check=ctx.check)
        return g.to(means.device), None, None


class MLPGBatch(Function):
    """Additive: :class:`MLPG` over a whole mini-batch -- ``(B, Tmax, D)`` zero-padded or flat
    ``(sum_T, D)`` means with per-utterance ``lengths`` -- one forward and one backward kernel launch
    for all utterances (the reference's ``MLPG`` is 2-D only and is looped over the batch).
    Per-frame variances of the same shape, or global ``(D,)``.  Returns float32."""

    @staticmethod
    def forward(ctx, means, variances, windows, lengths):
        assert means.dim() in (2, 3)
        ctx.windows = windows
        ctx.lengths = [int(n) for n in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
        ctx.save_for_backward(means, variances)
        device = _cuda_device(means)
        ctx.check = "deferred" if means.is_cuda else True
        y = G.mlpg_batch(means.detach().to(device), variances.detach().to(device), windows, lengths=ctx.lengths,
                         check=ctx.check)
        return y.to(torch.float32).to(means.device)

    @staticmethod
    def backward(ctx, grad_output):
        means, variances = ctx.saved_tensors
        device = _cuda_device(means)
        g = G.mlpg_grad_batch(variances.detach().to(device), ctx.windows, grad_output.detach().to(device), ctx.lengths,
                              check=ctx.check)
        return g.to(means.device), None, None, None


class UnitVarianceMLPG(Function):
    r"""MLPG for unit-variance inputs, ``y = R \mu`` (mlpg.py:70-172).

    ``f : (T x D) -> (T, static_dim)`` or ``f : (T*num_windows, static_dim) -> (T, static_dim)``,
    2-D or 3-D (batched) ``means``; ``R`` from :func:`nnmnkwii_b200.paramgen.unit_variance_mlpg_matrix`.
    """

    @staticmethod
    def forward(ctx, means, R):
        from . import _uvmlpg as uv

        ctx.save_for_backward(means, R)
        ctx.num_windows = R.shape[-1] // R.shape[0]
        T = R.shape[0]
        dim = means.dim()
        if dim == 2:
            T_, D = means.shape
            B = 1
            means3 = means.reshape(B, T_, D)
        else:
            B, T_, D = means.shape
            means3 = means
        reshaped = not (T == T_)  # mlpg.py:123: input already (T*nw, static_dim)?
        device = _cuda_device(means)
        band = uv.band_of(R, device)
        out = uv.apply_forward(band, means3.detach().to(device), reshaped).to(means.device)
        ctx.reshaped = reshaped
        if dim == 2:
            return out.view(-1, out.shape[-1])
        return out

    @staticmethod
    def backward(ctx, grad_output):
        from . import _uvmlpg as uv

        means, R = ctx.saved_tensors
        T = R.shape[0]
        dim = means.dim()
        if dim == 2:
            T_, D = means.shape
            B = 1
            grad_output = grad_output.reshape(B, T, -1)
        else:
            B, T_, D = means.shape
        device = _cuda_device(means)
        band = uv.band_of(R, device)
        grad = uv.apply_backward(band, grad_output.detach().to(device), ctx.reshaped, D).to(means.device)
        if dim == 2:
            return grad.view(-1, D), None
        return grad, None


def mlpg(means, variances, windows):
    """Maximum Likelihood Parameter Generation on tensors (mlpg.py:175-199).

    ``variances`` may be ``(T, D)`` or global ``(D,)`` (expanded over frames).
    """
    T, D = means.size()
    if not (variances.dim() == 1 and variances.shape[0] == D):  # a global (D,) variance stays 1-D
        assert means.size() == variances.size()
    return MLPG.apply(means, variances, windows)


def mlpg_batch(means, variances, windows, lengths):
    """Additive: batched :func:`mlpg` (see :class:`MLPGBatch`)."""
    return MLPGBatch.apply(means, variances, windows, lengths)


def unit_variance_mlpg(R, means):
    """Special case of MLPG assuming unit variances (mlpg.py:202-217).  NB argument order
    ``(R, means)`` here, ``(means, R)`` for ``UnitVarianceMLPG.apply`` -- as in the reference."""
    return UnitVarianceMLPG.apply(means, R)


__all__ = ["MLPG", "MLPGBatch", "UnitVarianceMLPG", "mlpg", "mlpg_batch", "unit_variance_mlpg"]
_ = np  # numpy is part of the reference module's namespace
